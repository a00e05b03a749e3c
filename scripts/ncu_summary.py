"""Turn the ncu captures of scripts/gpu_validate.sh (gpurun_out/) into the tracked summaries under profiles/:
  profiles/<tag>_launches_{async,sync}.txt   one update's launch list with per-kernel durations and family shares
  profiles/<tag>_gather_traffic.json         dram__bytes_{read,write}.sum per launch of the replay gather (bench.py `traffic`)
  profiles/<tag>_ncu_summary.txt             duration / DRAM bytes / tensor-pipe % / registers of the update's kernels
Usage: python scripts/ncu_summary.py [tag]   (runs here: ncu -i reads the reports without a GPU)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[2:]


def launches(mode):
    path = os.path.join(OUT, "launches_%s_%s.csv" % (mode, tag))
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    names = [(r[4], float(r[-1]) / 1e3) for r in rows]
    ends = [i for i, (n, _) in enumerate(names) if "nature_fused_opt" in n or "rmsprop_kernel" in n or "adam_kernel" in n]
    seg = names[ends[-2] + 1: ends[-1] + 1] if len(ends) >= 2 else names
    fam = collections.OrderedDict()
    for n, t in seg:
        k = n.split("(")[0].replace("void ", "").replace("b2rl::", "")
        k = k.split("<")[0] if k.startswith("at::") else k
        fam[k] = fam.get(k, 0.0) + t
    tot = sum(t for _, t in seg)
    with open(os.path.join(PROF, "%s_launches_%s.txt" % (tag, mode)), "w") as f:
        f.write("# ncu launch list of ONE eager DQN update (bench workload, B=512, %s replay), round-2 code\n" % mode)
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/profile_step.py --updates 2%s\n"
                % (" --replay sync" if mode == "sync" else ""))
        f.write("# durations are cold-cache and serialised by ncu: use the SHARES against bench.py's ms_per_step, not the sum\n# us      kernel\n")
        for n, t in seg:
            f.write("%8.1f  %s\n" % (t, n.replace("void ", "").replace("b2rl::", "")[:110]))
        f.write("# total %.1f us over %d launches\n#\n# share by kernel family\n" % (tot, len(seg)))
        for k, t in sorted(fam.items(), key=lambda kv: -kv[1]):
            f.write("#  %5.1f%%  %7.1f us  %s\n" % (100 * t / tot, t, k))
    print("wrote launches", mode, len(seg), "launches", round(tot, 1), "us")


def gather():
    rep = os.path.join(OUT, "prof_gather_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    hdr, rows = raw(rep)
    col = lambda n: hdr.index(n)
    r = rows[-1]
    f = lambda n: float(r[col(n)])
    unit = lambda n: {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[raw_units[col(n)]]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw_units = list(csv.reader(txt.splitlines()))[1]
    rd, wr = f("dram__bytes_read.sum") * unit("dram__bytes_read.sum"), f("dram__bytes_write.sum") * unit("dram__bytes_write.sum")
    d = dict(kernel=r[col("Kernel Name")], dram_bytes_read=int(rd), dram_bytes_write=int(wr), duration_us=f("gpu__time_duration.sum"),
             source="profiles/%s_gather_traffic.json: ncu --set full --clock-control none -k regex:gather_cvt (scripts/gpu_validate.sh), "
                    "dram__bytes_read.sum + dram__bytes_write.sum of one launch (200k-slot ring, batch 512)" % tag)
    json.dump(d, open(os.path.join(PROF, "%s_gather_traffic.json" % tag), "w"), indent=1)
    print("gather traffic", d)


def step():
    rep = os.path.join(OUT, "prof_step_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    hdr, rows = raw(rep)
    want = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tc_pipe%"), ("dram__bytes_read.sum", "dramRd"),
            ("dram__bytes_write.sum", "dramWr"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]
    with open(os.path.join(PROF, "%s_ncu_summary.txt" % tag), "w") as f:
        f.write("# ncu --set full --clock-control none --cache-control none --import-source on (scripts/gpu_validate.sh), one eager DQN update,\n"
                "# B=512, 84x84x4 uint8 frames, 200k-slot ring for the profile run (bench uses 1M slots; the kernels are identical)\n")
        f.write("# %-58s %s\n" % ("kernel", " ".join("%10s" % w[1] for w in want)))
        for r in rows:
            name = r[hdr.index("Kernel Name")].replace("void ", "")[:58]
            vals = []
            for k, _ in want:
                try:
                    vals.append("%10.2f" % float(r[hdr.index(k)]))
                except (ValueError, IndexError):
                    vals.append("%10s" % "-")
            f.write("  %-58s %s\n" % (name, " ".join(vals)))
    print("wrote ncu summary", len(rows), "kernels")


if __name__ == "__main__":
    os.makedirs(PROF, exist_ok=True)
    launches("async"), launches("sync"), gather(), step()
    for mode in ("async", "sync"):
        src = os.path.join(OUT, "timeline_%s_%s.txt" % (mode, tag))
        if os.path.exists(src):
            open(os.path.join(PROF, "%s_timeline_%s.txt" % (tag, mode)), "w").write(open(src).read())
    rc = os.path.join(OUT, "racecheck_%s.log" % tag)
    if os.path.exists(rc):
        lines = open(rc).read().splitlines()
        keep = [l for l in lines if "RACECHECK" in l or "passed" in l or "failed" in l or "error" in l.lower()][:60]
        open(os.path.join(PROF, "%s_racecheck.txt" % tag), "w").write(
            "# compute-sanitizer --tool racecheck on the sum-tree / replay parity tests (scripts/gpu_validate.sh)\n" + "\n".join(keep) + "\n")
