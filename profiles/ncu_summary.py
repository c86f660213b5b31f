#!/usr/bin/env python
"""Prints the per-kernel numbers quoted in profiles/*.md from an .ncu-rep (ncu -i <rep> --page raw --csv)."""
import csv, re, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_utchmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']

def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('----', re.sub(r'\(.*', '', r[ki])[-70:])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w} [{units[i]}] = {r[i]}")
        for i, h in enumerate(hdr):   # stall breakdown: top 4 reasons
            pass
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
        for v, h in sorted(stalls, reverse=True)[:4]:
            print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} = {v:.2f} warps/issue")

if __name__ == '__main__':
    main(sys.argv[1])
