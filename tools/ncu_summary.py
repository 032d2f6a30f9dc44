"""Print the metrics we care about from an .ncu-rep (run here on the CPU box: ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__cycles_elapsed.max',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for R in rows[2:]:
        d = dict(zip(H, R))
        print('kernel:', d.get('Kernel Name', '')[:60], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
        for k in WANT:
            if k in d: print(f'  {k:75s} {d[k]:>16s} {U[H.index(k)]}')
        for k in H:
            if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio') and 'average_warps' in k:
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v > 0.15: print(f'  stall {k.split("issue_stalled_")[1].split("_per_issue")[0]:40s} {v:8.2f}')
def traffic_json(out_path, landmarks, reps):
    """profiles/*_traffic.json for bench.py's roofline.traffic: DRAM bytes per filter of each kernel in the captures
    (one CTA per filter: per-launch bytes / grid size)."""
    import json
    kernels = {}
    for path in reps:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        H, U = rows[0], rows[1]
        for R in rows[2:]:
            d = dict(zip(H, R))
            import re
            mm = re.search(r'(k_\w+)', d.get('Kernel Name', ''))
            name = mm.group(1) if mm else d.get('Kernel Name', '')
            def val(k):
                v = float(d[k].replace(',', ''))
                u = U[H.index(k)].lower()
                return v * {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1.0)
            grid = int(d.get('launch__grid_size', '0').replace(',', '') or 0)
            if not grid:
                continue
            tot = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
            kernels[name] = {'dram_bytes_per_launch': tot, 'filters_per_launch': grid,
                             'dram_bytes_per_filter': tot / grid, 'capture': path.split('/')[-1]}
    with open(out_path, 'w') as f:
        json.dump({'landmarks': landmarks, 'source': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum',
                   'kernels': kernels}, f, indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--traffic-json':
        traffic_json(sys.argv[2], int(sys.argv[3]), sys.argv[4:])
    else:
        for p in sys.argv[1:]: main(p)
