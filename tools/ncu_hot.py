"""Summarise an ncu report: headline metrics + hottest SASS lines with their stall reasons.
   python tools/ncu_hot.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed',
        'smsp__mem_tensor_writes_op_stt.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
  if h in want:
    print(f'{h:100s} {u:16s} {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'SASS lines', len(data))
agg = {}
for r in data:
  for h in stalls:
    agg[h] = agg.get(h, 0) + int(r[ix[h]])
print('stall totals:', sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top_n]:
  s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
  best = sorted(s.items(), key=lambda kv: -kv[1])[:2]
  print(r[ix['# Samples']].rjust(7), r[ix['Instructions Executed']].rjust(10), r[ix['Source']][:72].ljust(72), best)
