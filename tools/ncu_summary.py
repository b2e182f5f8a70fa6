#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel the raw-page counters that the profiles/ notes quote,
and (--source) the stall-sample profile along the instruction stream in N buckets.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source 40] [--csv out.csv]"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__cluster_dim_x', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__inst_executed.sum']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    table = []
    for r in rows[2:]:
        rec = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[k] = (r[i], units[i])
        table.append(rec)
        print("----")
        for k, (v, u) in rec.items():
            print("%s = %s %s" % (k, v, u))
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(table))])
            for k in KEYS:
                if any(k in t for t in table):
                    w.writerow([k, next(t[k][1] for t in table if k in t)] + [t.get(k, ("", ""))[0] for t in table])
    if "--source" in sys.argv:
        B = int(sys.argv[sys.argv.index("--source") + 1])
        rows = page(rep, "source")
        ks, cur = [], None
        for r in rows:
            if r and r[0] == 'Kernel Name':
                cur = {'name': r[1], 'rows': []}
                ks.append(cur)
            elif r and r[0] == 'Address':
                cur['hdr'] = r
            elif cur is not None and r:
                cur['rows'].append(r)
        for k in ks:
            h = k['hdr']
            si, src = h.index('# Samples'), h.index('Source')
            tot = sum(int(r[si]) for r in k['rows']) or 1
            n = len(k['rows'])
            print(k['name'][:90], 'samples', tot, 'instructions', n)
            cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
            for b in range(B):
                seg = k['rows'][b * n // B:(b + 1) * n // B]
                s = sum(int(r[si]) for r in seg)
                st, ops = {}, {}
                for r in seg:
                    for i in cols:
                        v = int(r[i])
                        if v:
                            st[h[i]] = st.get(h[i], 0) + v
                    t = r[src].split()
                    op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
                    ops[op] = ops.get(op, 0) + 1
                print('%3d %5.1f%% %s | %s' % (b, 100.0 * s / tot, sorted(st.items(), key=lambda x: -x[1])[:3], sorted(ops.items(), key=lambda x: -x[1])[:4]))


if __name__ == "__main__":
    main()
