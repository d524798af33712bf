"""Warp-stall samples of an `ncu --set full --import-source on` report attributed to SOURCE LINES (needs a -lineinfo build).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass \
        --resolve-source-file gpt-st_b200/csrc/gproj2.cu,gpt-st_b200/csrc/mma_f16.cuh,gpt-st_b200/csrc/common.cuh > /tmp/src.csv
    python tools/ncu_source_stalls.py /tmp/src.csv [kernel-instance-index] [top-N]

Prints, for the chosen kernel instance of the report (default: the last), the share of stall samples per source line, the
instructions executed on that line and its three dominant stall reasons.  This is how profiles/ncu_gproj2_bwd_r01.md was made."""
import csv
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    secs, cur, i = [], None, 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            cur = {"file": r[1], "fn": rows[i + 1][1], "hdr": rows[i + 2], "rows": []}
            secs.append(cur)
            i += 3
            continue
        if cur is not None and r:
            cur["rows"].append(r)
        i += 1
    inst, seen, k = [], set(), []
    for s in secs:                       # the per-file sections of one kernel instance follow each other
        if s["file"] in seen:
            inst.append(k)
            k, seen = [], set()
        seen.add(s["file"])
        k.append(s)
    if k:
        inst.append(k)
    return inst


def analyse(k, top):
    tot, lines = 0, []
    for s in k:
        h = s["hdr"]
        ws, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        stall_cols = [(n, j) for j, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        for r in s["rows"]:
            if r[0] == "":               # SASS rows; the row with a line number carries the line's totals
                continue
            try:
                smp = int(r[ws])
            except ValueError:
                continue
            st = {n: int(r[j]) for n, j in stall_cols if r[j] not in ("", "-") and int(r[j]) > 0}
            lines.append((smp, s["file"].split("/")[-1], r[0], r[1].strip()[:100], int(r[ie]) if r[ie].isdigit() else 0, st))
            tot += smp
    lines.sort(reverse=True)
    print(f"{k[0]['fn'][:110]}\n{tot} stall samples")
    for smp, f, ln, src, n_inst, st in lines[:top]:
        top3 = ", ".join(f"{n[6:]} {v}" for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * smp / max(tot, 1):5.1f}%  {f}:{ln:>4}  inst {n_inst:>8}  {src}   [{top3}]")


if __name__ == "__main__":
    inst = sections(sys.argv[1])
    which = int(sys.argv[2]) if len(sys.argv) > 2 else len(inst) - 1
    print(f"{len(inst)} kernel instances in the report")
    analyse(inst[which], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
