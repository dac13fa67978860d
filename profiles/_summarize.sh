#!/bin/bash
# usage: profiles/_summarize.sh <report.ncu-rep> <out.txt> [top-lines]
# Text summary of one `ncu --set full` capture: key counters, stall reasons, per-source-line opcode mix.
set -e
rep=$1; out=$2; top=${3:-16}
tmp=$(mktemp -d)
ncu -i "$rep" --page raw --csv 2>/dev/null > $tmp/raw.csv
ncu -i "$rep" --page source --csv --print-source cuda,sass 2>/dev/null > $tmp/src.csv
python - "$tmp/raw.csv" > "$out" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in want:
        if w in d:
            print("%-72s %s %s" % (w, d[w], units[hdr.index(w)]))
    for h in hdr:
        if 'tensor' in h and 'pct_of_peak_sustained_active' in h and h not in want:
            print("%-72s %s %s" % (h, d[h], units[hdr.index(h)]))
    st = {}
    for h, v in zip(hdr, r):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                st[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(v.replace(',', ''))
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    print("warp stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    print()
PY
python "$(dirname "$0")/_perline.py" $tmp/src.csv $top >> "$out"
rm -rf $tmp
