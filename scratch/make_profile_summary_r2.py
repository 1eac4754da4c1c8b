"""Turn the round-2 ncu artefacts in gpurun_out/ into the committed summaries under profiles/."""
import csv, collections, json, subprocess, os, io
os.makedirs('profiles', exist_ok=True)
lines = [l for l in open('gpurun_out/r2_bench_launches.csv') if not l.startswith('==')]
open('profiles/r2_bench_launches.csv', 'w').writelines(lines)
def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name'].split('(')[0].replace('void ', '')[:60]
    agg.setdefault(k, []).append(float(row['Metric Value'].replace(',', '')) / 1e3)
def mean(k): return sum(agg[k]) / len(agg[k])
groups = collections.OrderedDict([
    ("the cfg2 fp32 loss step", lambda k: k.startswith('yb::fused_main_kernel<float, 4, 1')),
    ("the NMS step", lambda k: k.startswith('yb::nms_')),
    ("the task-aligned step", lambda k: k.startswith('yb::tal_')),
    ("the cfg5 bf16 loss step", lambda k: k.startswith('yb::fused_main_kernel<__nv_bfloat16')),
])
totals = {g: sum(mean(k) for k in agg if f(k)) for g, f in groups.items()}
out = ["# Round-2 profile summary (B200, sm_100a)", "",
       "## 1. ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 700` — per-launch times are cold-cache and serialised:",
       "compare SHARES, not absolutes (the CUDA-event numbers of `bench.py` are the timings). Raw list: `r2_bench_launches.csv`.",
       "The nearest-centre step is ONE launch now (`fused_main_kernel`: box, class, match roles and the reducer); the task-aligned",
       "step is four (`tal_decode_kernel`, `tal_gt_kernel`, `tal_cls_kernel`, `tal_finalize_kernel`; six in round 1).", "",
       "| kernel | launches | mean µs | share |", "|---|---|---|---|"]
for k, v in agg.items():
    m = mean(k)
    share = ("(the forward-only call of the bench's parity check)" if k.startswith('yb::fused_main_kernel<float, 4, 0') else
             "(the API leg's `loss.backward()`: returns at once, the upstream gradient is 1)" if k.startswith('yb::scale_kernel') else
             "(torch: host-side tensor preparation of the bench, not on the path)")
    for g, f in groups.items():
        if f(k): share = f"{100 * m / totals[g]:.1f} % of {g}"
    out.append(f"| `{k}` | {len(v)} | {m:.1f} | {share} |")
want = [('gpu__time_duration.sum', 'duration'), ('smsp__inst_executed.sum', 'warp instructions'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % (ncu peak)'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long-scoreboard stall / issue'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %')]
traffic, pipes = {}, {}
for title, rep, note in (("## 2. `ncu --set full` — `fused_main_kernel` at cfg2 fp32 (`scratch/prof_loss.py loss`)", 'gpurun_out/r2_loss_f32.ncu-rep', 'loss'),
                         ("## 3. `ncu --set full` — `fused_main_kernel` at cfg5 bf16 (`scratch/prof_loss.py loss_bf16_cfg5`)", 'gpurun_out/r2_loss_bf16_cfg5.ncu-rep', 'loss5'),
                         ("## 4. `ncu --set full` — NMS kernels at cfg4 (`scratch/prof_loss.py nms`)", 'gpurun_out/r2_nms.ncu-rep', 'nms'),
                         ("## 5. `ncu --set full` — task-aligned variant at cfg2 (`scratch/tal_once.py`)", 'gpurun_out/r2_tal_g.ncu-rep', 'tal')):
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["", title, "", "| kernel | " + " | ".join(n for _, n in want) + " |", "|---|" + "---|" * len(want)]
    for r in rows:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
        vals = [(r[idx[m]] + ' ' + units[idx[m]]).strip() if m in idx else 'n/a' for m, _ in want]
        out.append(f"| `{name}` | " + " | ".join(vals) + " |")
        def mb(m):
            v = float(r[idx[m]].replace(',', '')); u = units[idx[m]]
            return v * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1}[u]
        if note in ('loss', 'loss5', 'tal'):
            traffic[name if note != 'loss' else 'fused_main_kernel'] = int(mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'))
        if note == 'nms':
            pipes[name] = {"fma_pipe_pct": float(r[idx['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']]),
                           "alu_pipe_pct": float(r[idx['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']]),
                           "issue_active_pct": float(r[idx['smsp__issue_active.avg.pct_of_peak_sustained_active']]),
                           "duration_us": float(r[idx['gpu__time_duration.sum']])}
json.dump(traffic, open('profiles/r2_traffic.json', 'w'), indent=1)
json.dump(pipes, open('profiles/r2_nms_pipes.json', 'w'), indent=1)
notes = open('scratch/r2_profile_notes.md').read() if os.path.exists('scratch/r2_profile_notes.md') else ''
open('profiles/r2_summary.md', 'w').write("\n".join(out) + "\n\n" + notes)
print("\n".join(out[:30]))
