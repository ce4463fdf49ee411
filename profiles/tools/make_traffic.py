"""profiles/disagg_traffic.json from an `ncu --set full` capture of one interior msb_disaggregate call:
DRAM bytes (read + write) of its kernels per cell-timestep, stamped with the hash of the sources the
captured library was built from (bench.py refuses the file when the hash differs from the build).

    python profiles/tools/make_traffic.py rep.ncu-rep <cells> <steps of the captured call> <source hash> <capture name>
"""
import csv, json, os, subprocess, sys
rep, cells, steps, src_hash, cap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
kern = {}
for d in rows[2:]:
    name = d[ki].split('(')[0].replace('void ', '').split('<')[0]
    if name not in ('thermo_knot_kernel', 'sw_day_kernel', 'prec_day_kernel') or name in kern:
        continue                      # the first captured launch of each interior kernel
    kern[name] = {'read': float(d[ri]) * scale[units[ri]], 'write': float(d[wi]) * scale[units[wi]]}
total = sum(v['read'] + v['write'] for v in kern.values())
cts = cells * steps
out = {"what": f"DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, capture {cap}) of the interior "
               f"kernels of one msb_disaggregate call on CONUS 1/16 deg: {cells} cells x {steps} steps",
       "source_hash": src_hash, "kernels": kern, "cell_timesteps": cts, "dram_bytes": total,
       "dram_bytes_per_cell_timestep": total / cts, "algorithmic_bytes_per_cell_timestep": 33.333}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'disagg_traffic.json')
json.dump(out, open(path, 'w'), indent=1)
print(json.dumps(out, indent=1))
