// stands in for disagg_streams.cu in the mixed-kernel experiment (its code is part of mixed_interior.cu)
