"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/ (round 1)."""
import csv, collections, json, subprocess, os
os.makedirs('profiles', exist_ok=True)
subprocess.run('cp gpurun_out/r1_launches.csv profiles/r1_bench_launches.csv', shell=True, check=True)
def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]
lines = [l for l in open('profiles/r1_bench_launches.csv') if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name'].split('(')[0].replace('void ', '')[:60]
    agg.setdefault(k, []).append(float(row['Metric Value'].replace(',', '')) / 1e3)
def mean(k): return sum(agg[k]) / len(agg[k])
groups = collections.OrderedDict([
    ("the cfg2 fp32 loss step", lambda k: k.startswith('yb::fused_main_kernel<float') or k.startswith('yb::match_kernel<float') or k.startswith('yb::finalize_kernel')),
    ("the NMS step", lambda k: k.startswith('yb::nms_')),
    ("the task-aligned step", lambda k: k.startswith('yb::tal_')),
    ("the cfg5 bf16 loss step (+ finalize_kernel above)", lambda k: k.startswith('yb::fused_main_kernel<__nv_bfloat16') or k.startswith('yb::match_kernel<__nv_bfloat16')),
])
totals = {g: sum(mean(k) for k in agg if f(k)) for g, f in groups.items()}
out = ["# Round-1 profile summary (B200, sm_100a)", "",
       "## 1. ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` — per-launch times are cold-cache and serialised:",
       "compare SHARES, not absolutes (the CUDA-event numbers of `bench.py` are the timings). Raw list: `r1_bench_launches.csv`.", "",
       "| kernel | launches | mean µs | share |", "|---|---|---|---|"]
for k, v in agg.items():
    m = mean(k)
    share = "(torch: host-side tensor prep of the bench)"
    for g, f in groups.items():
        if f(k): share = f"{100 * m / totals[g]:.1f} % of {g}"
    out.append(f"| `{k}` | {len(v)} | {m:.1f} | {share} |")
want = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % (ncu peak)'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe %'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %')]
traffic = {}
for title, rep, note in (("## 2. `ncu --set full` — loss kernels at cfg2 (`scratch/prof_loss.py loss`)", 'gpurun_out/r1_loss_full.ncu-rep', 'loss'),
                         ("## 3. `ncu --set full` — NMS kernels at cfg4 (`scratch/prof_loss.py nms`)", 'gpurun_out/r1_nms_full.ncu-rep', 'nms'),
                         ("## 4. `ncu --set full` — task-aligned variant at cfg2 (`scratch/prof_loss.py tal`)", 'gpurun_out/r1_tal_full.ncu-rep', 'tal')):
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["", title, "", "| kernel | " + " | ".join(n for _, n in want) + " |", "|---|" + "---|" * len(want)]
    for r in rows:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
        vals = [(r[idx[m]] + ' ' + units[idx[m]]).strip() if m in idx else 'n/a' for m, _ in want]
        out.append(f"| `{name}` | " + " | ".join(vals) + " |")
        if note == 'loss':
            def b(m):
                return float(r[idx[m]].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[idx[m]]]
            traffic[name.split('<')[0]] = int(b('dram__bytes_read.sum') + b('dram__bytes_write.sum'))
out += ["", "## 5. Reading", "",
        f"* DRAM traffic per launch (read + write): `{json.dumps(traffic)}`; algorithmic bytes of `fused_main_kernel` at cfg2 = 1 238 630 400",
        "  (every head-output byte read once, every gradient byte written once) — traffic/algorithmic ≈ 0.97: nothing is re-read; the",
        "  shortfall is gradient lines still in L2 when the kernel ends.",
        "* The box role's (GT, tile) pruning shows in the cfg5 line of section 1: `fused_main_kernel<__nv_bfloat16, 8, 1>` fell from",
        "  235 us to 165 us per launch when the pruning and the coarse-tiles-first launch order went in (same bytes).",
        "* Task-aligned variant: `tal_candidates_kernel` is issue-bound (53 % issue-active at 40 % occupancy, 297 MB of DRAM reads in",
        "  140 us), `tal_fg_kernel` moves 174 MB of 64-byte bursts for 139 MB of useful 32-byte sectors, `tal_cls_kernel` streams",
        "  928 MB at 6.1 TB/s (152 us, eight short CTAs per tile) with every gradient store coalesced (the foreground rows are merged in, not scattered); its issue",
        "  activity fell from 67 % to 36 % with the branch-free packed softplus/sigmoid (ALU pipe 49 % -> 25 %).",
        "* No tensor-pipe activity anywhere (nothing on this path is a dense contraction).",
        "* Blackwell-specific SASS: `FFMA2` / `FMUL2` / `FADD2` (packed FP32, PTX `fma.rn.f32x2`) in `fused_main_kernel`:",
        "  `cuobjdump -sass custom-yolo-implmentation_b200/csrc/libyolo_boxpath.so | grep -c FFMA2`.",
        "* compute-sanitizer is closed on this pool (gpurun refuses it), so memory safety rests on the parity suite on ragged /",
        "  unaligned shapes (A = 189, empty images, > 128 GT per image, > 9216 NMS candidates)."]
open('profiles/r1_summary.md', 'w').write("\n".join(out) + "\n")
json.dump(traffic, open('profiles/r1_traffic.json', 'w'), indent=1)
print("\n".join(out))
