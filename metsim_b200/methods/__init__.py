"""Method modules in the shape the reference registry expects (metsim/methods/)."""
