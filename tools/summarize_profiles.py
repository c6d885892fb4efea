"""Turns gpurun_out/launches_*.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum --csv) and .ncu-rep captures
into the tracked summaries under profiles/.  Runs here (no GPU): `python tools/summarize_profiles.py <tag>`."""
import csv, json, re, subprocess, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, prof = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    L = {}
    for r in rows[hi + 2:]:
        if len(r) <= iv: continue
        d = L.setdefault(int(r[iid]), {"k": r[ik]})
        d[r[im]] = float(r[iv].replace(",", ""))
    return [L[i] for i in sorted(L)]

def short(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"<unnamed>::", "", k)
    m = re.match(r"([\w:]+(?:<[^(]*?>)?)", k)
    return (m.group(1) if m else k)[:70]

src = os.path.join(go, "launches_%s.csv" % tag)
if os.path.exists(src):
    L = launches(src)
    n = len(L)
    step = L[n // 2:]                               # second half = the timed step (first half = its warm-up twin)
    tot = sum(d["gpu__time_duration.sum"] for d in step)
    agg = {}
    for d in step:
        a = agg.setdefault(short(d["k"]), [0, 0.0, 0.0])
        a[0] += 1; a[1] += d["gpu__time_duration.sum"]
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    with open(os.path.join(prof, "%s_launches_bench_step.csv" % tag), "w") as f:
        f.write("id,kernel,gpu__time_duration_us,dram_read_bytes,dram_write_bytes\n")
        for i, d in enumerate(step):
            f.write('%d,"%s",%.2f,%d,%d\n' % (i, short(d["k"]), d["gpu__time_duration.sum"] / 1e3,
                                               d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)))
    conv = [d for d in step if "conv_stream" in d["k"]]
    conv_t = sum(d["gpu__time_duration.sum"] for d in conv)
    conv_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in conv)
    with open(os.path.join(prof, "%s_launches_summary.md" % tag), "w") as f:
        f.write("# %s -- ncu launch list of one `bench.py` step\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                "--csv --log-file gpurun_out/launches_%s.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e` "
                "(second of the two captured steps; 8 images x 5 pyramid levels x {plain, mirrored}, batched 16 per level). "
                "Serialised, cold-cache durations: read SHARES. Per-launch rows: `%s_launches_bench_step.csv`.\n\n" % (tag, tag))
        f.write("| kernel | launches | total ms | share | avg us | DRAM GB (rd+wr) |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.2f | %.1f%% | %.1f | %.2f |\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e3, a[2] / 1e9))
        f.write("\nTotal %.1f ms GPU time for the step. tcgen05 conv (`conv_stream_kernel`) share: **%.1f%%**; "
                "DRAM traffic of its %d launches: %.1f GB = **%.1f MB per launch** on average.\n"
                % (tot / 1e6, 100 * conv_t / tot, len(conv), conv_b / 1e9, conv_b / len(conv) / 1e6))
    json.dump({"conv_launches": len(conv), "conv_dram_bytes_per_launch": conv_b / len(conv), "conv_share_of_step": conv_t / tot,
               "source": "profiles/%s_launches_summary.md (ncu launch list, one bench step)" % tag},
              open(os.path.join(prof, "%s_conv_traffic.json" % tag), "w"), indent=1)
    print(open(os.path.join(prof, "%s_launches_summary.md" % tag)).read())

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__cluster_size", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
out = []
for rep in sorted(os.listdir(go)):
    if not (rep.startswith(tag + "_") and rep.endswith(".ncu-rep")): continue
    r = subprocess.run(["ncu", "-i", os.path.join(go, rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3: continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out.append((rep, d))
if out:
    with open(os.path.join(prof, "%s_ncu_full.md" % tag), "w") as f:
        f.write("# %s -- `ncu --set full --clock-control none --import-source on` captures (one launch each)\n\n" % tag)
        f.write("Commands: `ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 2 -c 1 -o gpurun_out/<name> "
                "python tools/ncu_one.py 8 <Cin> <Cout> <H> <W> 3 1 <fmt>` on shapes of the 2048-px level (conv4_2: 512->512 at 256x256; "
                "conv2_2: 128->128 at 1024x1024; conv1_2: 64->64 at 2048x2048; fmt 0 = split fp16, 1 = fp16 + fp8) and the same for "
                "`conv1_tc` (3->64 at 2048x2048; captured inside tools/level_conv_only.py) and the post-processing kernels (first five "
                "launches of smoke()).\n\n")
        f.write("| capture | " + " | ".join(k.replace("_", "\\_") for k in KEYS) + " |\n|---|" + "---:|" * len(KEYS) + "\n")
        for rep, d in out:
            f.write("| %s | " % rep + " | ".join("%s %s" % (d.get(k, ("", ""))[0], d.get(k, ("", ""))[1]) for k in KEYS) + " |\n")
    print(open(os.path.join(prof, "%s_ncu_full.md" % tag)).read()[:3000])

# ---- BASELINE configs[1]: per-conv ncu metrics of one whole 2048x2048 level (tools/level_conv_only.py 2048 1 under ncu) ----
lv = os.path.join(go, "%s_level2048_ncu.csv" % tag)
if os.path.exists(lv):
    L = launches(lv)
    names = ["conv1_2+pool", "conv2_1", "conv2_2+pool", "conv3_1", "conv3_2", "conv3_3+pool", "conv4_1", "conv4_2", "conv4_3+pool",
             "conv5_1", "conv5_2", "conv5_3", "conv5_256", "conv4_256", "conv4_fuse_final", "conv4_fuse_final_dim_red", "head_1",
             "head_2", "head_4"]
    n = len(names)
    if len(L) == 4 * n:                             # per format: one warm-up pass + one timed pass
        pipe = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
        with open(os.path.join(prof, "%s_level2048_ncu.md" % tag), "w") as f:
            f.write("# %s -- every tcgen05 conv launch of one 2048x2048 pyramid level under ncu (BASELINE configs[1])\n\n" % tag)
            f.write("Command: `ncu --metrics gpu__time_duration.sum,%s,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                    "-k regex:conv_stream --csv python tools/level_conv_only.py 2048 1` (dilated-head net, batch 1; second pass of each "
                    "operand format).  Durations are cold-cache / serialised; the CUDA-event table of the same launches is "
                    "`%s_level2048_conv_only.txt`.\n\n" % (pipe, tag))
            for title, off in (("h2 (split fp16, 3 MMAs per 16 channels)", n), ("hf8 (f16 + f8 correction, 2 MMAs per 16 channels)", 3 * n)):
                f.write("## operand format %s\n\n| layer | us | tensor pipe active %% | DRAM MB read | DRAM MB written |\n|---|---:|---:|---:|---:|\n" % title)
                tot = 0.0
                for name, d in zip(names, L[off:off + n]):
                    tot += d["gpu__time_duration.sum"]
                    f.write("| %s | %.1f | %.1f | %.1f | %.1f |\n" % (name, d["gpu__time_duration.sum"] / 1e3, d.get(pipe, float("nan")),
                                                                    d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
                f.write("| **all 19** | **%.1f** | | | |\n\n" % (tot / 1e3))
        print(open(os.path.join(prof, "%s_level2048_ncu.md" % tag)).read())
    else:
        print("level2048 csv: %d launches, expected %d" % (len(L), 4 * n))
