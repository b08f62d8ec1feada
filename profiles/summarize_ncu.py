#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md quotes.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep 3.6e8 > profiles/rNN_ncu_<what>_summary.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
       'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
       'smsp__inst_executed.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
       'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
       'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
       'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum']


def ncu(rep, page):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout


def main():
    rep, npts = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, 'raw'))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    for r in rows[2:]:
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"## {r[ix['Kernel Name']]}\n\n| metric | value |\n|---|---|")
        for m in RAW:
            if m in ix:
                print(f"| `{m}` | {r[ix[m]]} {units[ix[m]]} |")
        try:
            rd, wr = float(r[ix['dram__bytes_read.sum']]), float(r[ix['dram__bytes_write.sum']])
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
            tr = rd * scale[units[ix['dram__bytes_read.sum']]] + wr * scale[units[ix['dram__bytes_write.sum']]]
            print(f"| **DRAM traffic per launch** | {tr / 1e9:.3f} GB |")
            if npts:
                print(f"| DRAM bytes per point | {tr / npts:.2f} |")
                print(f"| warp-instructions per point x32 | {float(r[ix['smsp__inst_executed.sum']]) * 32 / npts:.1f} |")
            gl, gs = float(r[ix['l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']]), float(r[ix['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']])
            sl, ss = float(r[ix['l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']]), float(r[ix['l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']])
            print(f"| sectors / request, global loads (LSU path; point data arrives by TMA) | {gs / max(gl, 1):.2f} |")
            print(f"| sectors / request, global stores | {ss / max(sl, 1):.2f} |")
        except Exception as e:                      # noqa: BLE001
            print(f"| (derived metrics failed: {e}) | |")
        print()
    src = list(csv.reader(io.StringIO(ncu(rep, 'source'))))
    secs, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            secs.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for sec in secs[:1]:
        hdr = sec['rows'][0]
        data = [r for r in sec['rows'][1:] if len(r) == len(hdr) and r[0] != 'Address']
        ix = {h: i for i, h in enumerate(hdr)}

        def I(r, h):
            try:
                return int(r[ix[h]])
            except Exception:                       # noqa: BLE001
                return 0
        tot = sum(I(r, '# Samples') for r in data)
        print(f"## warp-stall samples ({tot} samples) and SASS mix — {sec['name']}\n\n| stall | share |\n|---|---|")
        st = {h: sum(I(r, h) for r in data) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
        for h, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
            print(f"| {h} | {100 * v / max(tot, 1):.1f} % |")
        op = collections.Counter()
        for r in data:
            m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[ix['Source']].strip())
            op[m.group(2) if m else '?'] += I(r, 'Instructions Executed')
        print("\n| SASS op | thread-instr per point |\n|---|---|" if npts else "\n| SASS op | warp-instr |\n|---|---|")
        for k, v in op.most_common(16):
            print(f"| {k} | {v * 32 / npts:.2f} |" if npts else f"| {k} | {v} |")
        proof = [k for k in op if k.startswith(('UBLKCP', 'SYNCS', 'UTMA'))]
        print(f"\nBlackwell/Hopper async-copy evidence in SASS: {', '.join(sorted(proof)) or 'none'}")


if __name__ == '__main__':
    main()
