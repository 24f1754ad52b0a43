"""Summarise an ncu report: key raw metrics per launch + a per-SASS-line table
(instruction share, average active lanes, stall-sample share).
    python tools/ncu_summary.py raw.csv [source.csv]"""
import csv, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size',
 'smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:60])
    for w in WANT:
        if w in hdr: print(f"  {w:85s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]].replace(',', ''))
        except Exception: return 0.0
    data = [r for r in rows[2:] if len(r) > ix['# Samples'] and r[0] != 'Address']
    ti = sum(f(r, 'Instructions Executed') for r in data); tt = sum(f(r, 'Thread Instructions Executed') for r in data); ts = sum(f(r, '# Samples') for r in data)
    print(f"SASS lines {len(data)}  warp-inst {ti:.3e}  thread-inst {tt:.3e}  avg lanes {tt/ti:.2f}  samples {ts:.0f}")
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    for n, r in enumerate(data):
        ie, te, sm = f(r, 'Instructions Executed'), f(r, 'Thread Instructions Executed'), f(r, '# Samples')
        if ie / ti > thr or sm / ts > thr * 1.5:
            top = sorted(((f(r, k), k) for k in stalls), reverse=True)[:2]
            tops = ' '.join(f"{k[6:]}={v:.0f}" for v, k in top if v > 0)
            print(f"{n:4d} inst {100*ie/ti:5.2f}% lanes {te/max(ie,1):5.1f} samp {100*sm/ts:5.2f}%  {r[ix['Source']][:70]:70s} {tops}")
