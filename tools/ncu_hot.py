"""Summarise an ncu report: headline metrics + hottest SASS lines with their stall reasons.
   python tools/ncu_hot.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active_realtime.avg.pct', 'sm__pipe_tensor_subpipe_hmma_cycles_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_uniform', 'lts__throughput.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'smsp__issue_active.avg.pct', 'sm__cycles_elapsed.max', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__sass_inst_executed_op_shared']
for h, u, v in zip(hdr, units, vals):
  if any(h.startswith(w) or w in h for w in want) and 'TriageCompute' not in h or 'tensor' in h and 'pct' in h:
    print(f'{h:90s} {u:12s} {v}')
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
