"""Digest of the round's .ncu-rep captures (gpurun_out/) into tracked files under profiles/:

    profiles/r2_ncu_<name>.txt     per-launch metrics of one capture (the ncu_summary.py key list)
    profiles/ncu_summary.json      what bench.py may quote: per kernel family, the launch count captured, mean duration,
                                   DRAM bytes per launch (read + write), fmaheavy pipe activity, registers, and where it
                                   came from.  bench.py never carries such figures as literals.

usage: python tools/ncu_digest.py <family>=<file.ncu-rep>[:note] ...   (families: msm_accumulate, ntt_pass, evaluate_h, ...)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main(argv):
    out_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    summary = {}
    if os.path.exists(out_json):
        summary = json.load(open(out_json))
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for arg in argv:
        fam, rest = arg.split("=", 1)
        rep, _, note = rest.partition(":")
        name = os.path.splitext(os.path.basename(rep))[0]
        txt = os.path.join(ROOT, "profiles", "r2_ncu_%s.txt" % name.replace("r2_", ""))
        ncu_summary.main(rep, txt)
        rows = raw_rows(rep)
        skip = int(os.environ.get("ZK_DIGEST_SKIP", "0"))          # launches at the head of the capture that belong to another step
        rows = rows[skip:]
        if not rows:
            print("no launches in", rep)
            continue
        dur = [num(r.get("gpu__time_duration.sum")) for r in rows]
        unit_scale = 1.0
        rd = [num(r.get("dram__bytes_read.sum")) for r in rows]
        wr = [num(r.get("dram__bytes_write.sum")) for r in rows]
        # ncu prints bytes in a per-column unit (row 2 of the csv); re-read the units
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        units = dict(zip(rr[0], rr[1]))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        time_scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
        s_rd, s_wr = scale.get(units.get("dram__bytes_read.sum"), 1.0), scale.get(units.get("dram__bytes_write.sum"), 1.0)
        s_t = time_scale.get(units.get("gpu__time_duration.sum"), 1.0)
        tot_bytes = sum((a or 0) * s_rd + (b or 0) * s_wr for a, b in zip(rd, wr))
        tot_us = sum((d or 0) * s_t for d in dur)
        fma = [num(r.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")) for r in rows]
        w = [(d or 0) * s_t for d in dur]
        fma_w = sum(f * x for f, x in zip(fma, w) if f is not None) / max(1e-9, sum(x for f, x in zip(fma, w) if f is not None))
        summary[fam] = {
            "source": "%s: ncu --set full --clock-control none, %d launch(es) of %s; kernels of commit %s%s"
                      % (os.path.relpath(txt, ROOT), len(rows), rows[0].get("Kernel Name", "?")[:60], head, ("; " + note) if note else ""),
            "launches_captured": len(rows),
            "mean_duration_us": tot_us / len(rows),
            "dram_bytes_per_launch": tot_bytes / len(rows),
            "dram_bytes_total": tot_bytes,
            "fmaheavy_pipe_active_pct": fma_w,
            "registers_per_thread": num(rows[0].get("launch__registers_per_thread")),
            "k": 19,
        }
        print(fam, json.dumps(summary[fam])[:300])
    with open(out_json, "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
