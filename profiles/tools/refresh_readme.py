"""Rewrite the measured numbers of profiles/README.md's round-2 tables from the committed evidence files
(profiles/r02/*.json, r02_launch_list_summary.txt), so that the prose never lags the JSON."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(ROOT)
last = lambda f: json.loads(open(f).read().strip().splitlines()[-1])
d, r, c = last('profiles/r02/bench_n1.json'), last('profiles/r02/bench_reference.json'), last('profiles/r02/bench_c2_global_half_deg.json')
p = 'profiles/README.md'; s = open(p).read()
br, e = d['breakdown_ms'], d['e2e']
s = re.sub(r"(\| `value` \(inputs resident in HBM\) \| 6\.66e10 cell-timesteps/s, 63\.5 ms per CONUS year \| )\*\*.*?\*\*.*? \|",
           r"\g<1>**%.2fe10 cell-timesteps/s, %.1f ms** (power-capped at 990-1000 W, SM clock %d MHz median) |" % (d['value'] / 1e10, d['ms_per_step'], d['clocks']['sm_mhz']), s)
s = re.sub(r"(\| solar tables / daily MTCLIM / disaggregation \(CUDA events\) \| 7\.9 / 13\.2 / 42\.6 ms \| ).*? ms \|",
           r"\g<1>%.1f / %.1f / %.1f ms |" % (br['solar_tables'], br['daily_mtclim'], br['disaggregate']), s)
s = re.sub(r"(\| disaggregation stage, fraction of the measured 6457 GB/s \(33\.33 B per cell-timestep\) \| 0\.513 \| )[0-9.]+( \|)", r"\g<1>%.3f\2" % d['roofline']['frac'], s)
s = re.sub(r"(\| whole step, same definition \(SURVEY 8d\) \| 0\.344 \| )[0-9.]+( \|)", r"\g<1>%.3f\2" % d['roofline']['whole_step_frac'], s)
s = re.sub(r"(\| `e2e` \(host buffers through the C ABI, copies inside the timer\) \| 1\.68e9, no consumer \| ).*? \|\n",
           r"\g<1>%.2fe9 (%.2f s per step, %.2f of the ceiling) with every tile consumed by the native sink (%.2f s copies only); D2H ceiling of the box %.1f GB/s |\n"
           % (e['value'] / 1e9, e['ms_per_step'] / 1e3, e['fraction_of_d2h_ceiling'], e['ms_per_step_copies_only'] / 1e3, e['d2h_ceiling_gbs_all_ranks']), s)
s = re.sub(r"(\| reference arm \(same box, 16 host cores, 256 cells\) \| 1\.82e6 \| ).*? \|",
           r"\g<1>%.2fe6 in the bench line, %.2fe6 in the `--impl reference` run |" % (d['cpu_baseline']['value'] / 1e6, r['value'] / 1e6), s)
txt = open('profiles/r02_launch_list_summary.txt').read()
rows = [l for l in txt.splitlines() if 'launches' in l]
tot = float(txt.split('total')[1].split()[0])
names = {'thermo_knot_kernel': '`thermo_knot_kernel` (temp, vapour pressure, rel. humidity, air pressure, longwave, wind; interior days; 97-day calls)',
         'prec_day_kernel': '`prec_day_kernel` (precipitation; interior days)', 'sw_day_kernel': '`sw_day_kernel` (shortwave; interior days)',
         'daily_fused_kernel': '`daily_fused_kernel` (tdew fixed point in one pass, candidate planes)', 'solar_moments_kernel': '`solar_moments_kernel`',
         'solar_rows_kernel': '`solar_rows_kernel`'}
tab = ['| Kernel | ms per step (5 steps in the list) | share | ms per launch |', '|---|---|---|---|']
other = 0.0
for l in rows:
    t = l.split(); ms = float(t[0]); avg = float(t[6]); name = ' '.join(t[8:])
    key = [k for k in names if k in name]
    if key and 'reduce' not in name:
        tab.append(f'| {names[key[0]]} | {ms / 5:.2f} | {100 * ms / tot:.1f} % | {avg:.2f} |')
    else:
        other += ms
tab.append(f'| edge days (`disagg_prepare_kernel` + `disagg_kernel<thermo / streams>`), reduce kernels, the replay over an empty list | {other / 5:.2f} | {100 * other / tot:.1f} % | |')
tab.append(f'| total of the list | {tot / 5:.2f} | | |')
i = s.index('| Kernel | ms per step (5 steps in the list)'); j = s.index('(ncu times each launch alone at full clocks')
s = s[:i] + '\n'.join(tab) + '\n\n' + s[j:]
s = re.sub(r"(the bench step, power-capped at 990 W with the SM clock at 1845-1950 MHz, takes )[0-9.]+( ms)", r"\g<1>%.1f\2" % d['ms_per_step'], s)
s = re.sub(r"\| patches of 4 rows \+ window-major copy of the shortwave table \(final code, `r02/bench_c2_global_half_deg.json`\) \| .*? \|\n",
           "| patches of 4 rows + window-major copy of the shortwave table (final code, `r02/bench_c2_global_half_deg.json`) | **%.2fe10** | %.1f (%.1f / %.1f / %.1f) |\n"
           % (c['value'] / 1e10, c['ms_per_step'], c['breakdown_ms']['solar_tables'], c['breakdown_ms']['daily_mtclim'], c['breakdown_ms']['disaggregate']), s)
open(p, 'w').write(s)
print("value %.3e  %.2f ms  e2e %.3e" % (d['value'], d['ms_per_step'], e['value']))
