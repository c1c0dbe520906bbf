"""From the condensed `--set full` capture of one pipelined step (tools/ncu_summary.py output) -> the stage-2 k4 block of ONE forward:
DRAM bytes (bench.py's roofline.traffic reads profiles/r02_stage2_traffic.json), L2->SM bytes, tensor-pipe activity, per layer.
usage: python tools/ncu_stage2_traffic.py profiles/r02_ncu_full_one_step.csv profiles/r02_stage2_traffic.json profiles/r02_ncu_stage2_layers.txt"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
col = lambda name: next(i for i, c in enumerate(h) if c.startswith(name))
K, G, D, DR, DW, L2, TP, TS = col('kernel'), col('grid'), col('duration'), col('dram_read'), col('dram_write'), col('l2_to_sm_read'), col('tensor_pipe_pct'), col('tensor_subpipe')
unit = lambda i: h[i].split('[')[1].rstrip(']') if '[' in h[i] else ''
scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'Tbyte': 1e12}
names = [r[K] for r in rows[1:]]
start = next(i for i, n in enumerate(names) if 'k_conv3x3_cin1' in n) + 1
end = next(i for i in range(start, len(names)) if 'k_conv3x3_cout1' in names[i])
blk = rows[1 + start:1 + end]
layer_names = ['c1', 'c2', 'c3', 'c4', 'c5', 'c6', 'c7', 'd0', 'd1', 'd2', 'd3', 'd4', 'd5', 'd6']
out, li = [], -1
tot = dict(dur=0.0, dram=0.0, l2=0.0, tp=0.0)
for r in blk:
    conv = 'k_conv_tc' in r[K] or 'k_conv_halo' in r[K]
    if conv:
        li += 1
    dur = float(r[D]); dram = float(r[DR]) * scale[unit(DR)] + float(r[DW]) * scale[unit(DW)]; l2 = float(r[L2]) * scale[unit(L2)]
    tp = float(r[TP]) if r[TP] else 0.0
    tot['dur'] += dur; tot['dram'] += dram; tot['l2'] += l2; tot['tp'] += tp * dur
    out.append(f"{layer_names[li] if li < 14 else '?':3s} {'conv  ' if conv else 'reduce'} grid {r[G]:>5s}  {dur:6.1f} us  dram {dram / 1e6:7.2f} MB  l2->sm {l2 / 1e6:7.1f} MB ({l2 / dur / 1e6:5.2f} TB/s)  tensor pipe {tp:5.1f} %")
n_conv = sum(1 for r in blk if 'k_conv_tc' in r[K] or 'k_conv_halo' in r[K])
summary = (f"stage-2 k4 block of one 384x512 forward under ncu (serialised, cold L1, --clock-control none): {len(blk)} launches ({n_conv} conv + {len(blk) - n_conv} "
           f"split-K reduce), {tot['dur']:.1f} us, DRAM read+write {tot['dram'] / 1e6:.1f} MB, L2->SM {tot['l2'] / 1e6:.0f} MB, time-weighted tensor pipe {tot['tp'] / tot['dur']:.1f} %")
open(sys.argv[3], 'w').write(summary + '\n' + '\n'.join(out) + '\n')
json.dump(dict(dram_bytes_per_forward=tot['dram'], l2_to_sm_bytes_per_forward=tot['l2'], launches=len(blk), duration_us_under_ncu=tot['dur'],
               tensor_pipe_pct_time_weighted=tot['tp'] / tot['dur'],
               source=f'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the {len(blk)} launches of one stage-2 k4 block: {sys.argv[1]} (tools/gpu_profile_r02.sh, tools/ncu_stage2_traffic.py)'),
          open(sys.argv[2], 'w'), indent=1)
print(summary); print('\n'.join(out))
