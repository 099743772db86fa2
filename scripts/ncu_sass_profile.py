"""Where do a kernel's instructions and stall samples go?  Reads an `ncu --set full
--import-source on` report and prints, per window of consecutive SASS instructions, the warp
instructions executed per 32 particles, the share of the kernel, the stall samples and the
dominant opcodes; or lists single instructions with their long-scoreboard samples.

    python scripts/ncu_sass_profile.py gpurun_out/stage_r01.ncu-rep                 # kernels in the report
    python scripts/ncu_sass_profile.py REPORT --kernel "stage_kernel<(int)1" --window 100
    python scripts/ncu_sass_profile.py REPORT --kernel gather --list 0 120 --min-samples 100
"""
import argparse
import collections
import csv
import io
import subprocess


def sections(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"]:
            cur["rows"].append(r)
    seen, uniq = set(), []
    for s in secs:                      # the export repeats every kernel once per view
        if s["name"] not in seen:
            seen.add(s["name"])
            uniq.append(s)
    return uniq


def opcode(text):
    t = text.split()
    return t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="substring of the kernel name")
    ap.add_argument("--window", type=int, default=100)
    ap.add_argument("--particles", type=int, default=16_000_000, help="threads of the launch (one per particle)")
    ap.add_argument("--list", nargs=2, type=int, default=None, metavar=("FROM", "TO"))
    ap.add_argument("--min-samples", type=int, default=0)
    a = ap.parse_args()
    secs = sections(a.report)
    if a.kernel is None:
        for s in secs:
            h = s["hdr"]
            ex = sum(int(r[h.index("Instructions Executed")]) for r in s["rows"])
            print(f"{s['name']}: {len(s['rows'])} SASS instructions, {ex} warp instructions executed")
        return
    s = next(x for x in secs if a.kernel in x["name"])
    h, rows = s["hdr"], s["rows"]
    i_ex, i_s, i_l = h.index("Instructions Executed"), h.index("# Samples"), h.index("stall_long_sb")
    warps = a.particles / 32
    total = sum(int(r[i_ex]) for r in rows)
    if a.list:
        for k in range(a.list[0], min(a.list[1], len(rows))):
            r = rows[k]
            if int(r[i_s]) >= a.min_samples:
                print(f"{k:5d} ex/warp {int(r[i_ex]) / warps:5.2f} samples {r[i_s]:>5} long_sb {r[i_l]:>5}  {r[1].strip()[:90]}")
        return
    for st in range(0, len(rows), a.window):
        blk = rows[st:st + a.window]
        ex = sum(int(r[i_ex]) for r in blk)
        if ex / warps < 1:
            continue
        ops = collections.Counter(opcode(r[1]) for r in blk)
        top = " ".join(f"{o}:{c}" for o, c in ops.most_common(6))
        print(f"{st:5d}-{st + a.window:5d} ex/warp {ex / warps:6.1f} ({100 * ex / total:4.1f}%) "
              f"samples {sum(int(r[i_s]) for r in blk):5d} | {top}")


if __name__ == "__main__":
    main()
