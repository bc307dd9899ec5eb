"""Where a kernel spends its stall samples, from an ncu report with source (read here, no GPU):

    ncu -i gpurun_out/prof_step_r01v10.ncu-rep --page source --csv --print-source sass -k regex:gate_coarse > /tmp/k.csv
    python tools/ncu_buckets.py /tmp/k.csv [instructions per bucket, default 40]

Prints the stall reasons of the whole kernel and, per run of N SASS instructions, the share of
stall samples, executed instructions and shared-memory wavefronts together with the memory /
barrier / shuffle opcodes of the run (enough to recognise the phase of the kernel)."""
import csv
import sys
from collections import Counter

MARK = ("ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "LD.", "ST.", "TEX", "TLD", "REDUX", "SHFL", "VOTE", "BAR", "MUFU",
        "UTMALDG", "SYNCS", "WARPSYNC", "NANOSLEEP")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr_at = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_at]
    data = []
    for r in rows[hdr_at + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break  # a second launch of the kernel follows: keep the first
        data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0

    tot_s = sum(f(r, "# Samples") for r in data) or 1.0
    tot_i = sum(f(r, "Instructions Executed") for r in data) or 1.0
    tot_w = sum(f(r, "L1 Wavefronts Shared") for r in data) or 1.0
    print("%d SASS instructions, %d samples, %.3g warp instructions, %.3g shared-memory wavefronts" % (len(data), tot_s, tot_i, tot_w))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(f(r, h) for r in data), h) for h in stalls), reverse=True)[:8]
    print("stalls:", ", ".join("%s %.1f %%" % (h[6:], 100 * v / tot_s) for v, h in agg))
    for b in range(0, len(data), per):
        seg = data[b:b + per]
        s = sum(f(r, "# Samples") for r in seg)
        i = sum(f(r, "Instructions Executed") for r in seg)
        w = sum(f(r, "L1 Wavefronts Shared") for r in seg)
        if s / tot_s < 0.005 and i / tot_i < 0.005:
            continue
        ops = []
        for r in seg:
            t = r[ix["Source"]].split()
            if t:
                ops.append(t[1] if t[0].startswith("@") and len(t) > 1 else t[0])
        marks = Counter(o for o in ops if o.startswith(MARK))
        top = max(seg, key=lambda r: f(r, "# Samples"))
        print("%5d  samples %5.1f %%  instr %5.1f %%  smem wavefronts %5.1f %%  %s | top %s (%.1f %%)" % (
            b, 100 * s / tot_s, 100 * i / tot_i, 100 * w / tot_w, dict(marks), " ".join(top[ix["Source"]].split())[:48], 100 * f(top, "# Samples") / tot_s))


if __name__ == "__main__":
    main()
