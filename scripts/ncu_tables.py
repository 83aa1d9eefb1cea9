"""Turn the ncu outputs under gpurun_out/ into the markdown tables of profiles/*.md.
   python scripts/ncu_tables.py launches <launch-list.csv>        (per-kernel totals of the LAST proof in the capture)
   python scripts/ncu_tables.py full <raw-page.csv>               (one column per captured launch; make the csv with
                                                                  ncu -i X.ncu-rep --page raw --csv > raw.csv)"""
import collections, csv, re, sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, mi, ii = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("ID")
    L = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        d = L.setdefault(r[ii], {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "")})
        d[r[mi]] = float(r[vi].replace(",", ""))
    ls = list(L.values())
    last_expand = max(i for i, d in enumerate(ls) if "expand_bases" in d["name"])      # end of the proving-key load
    proof = ls[last_expand + 1:]
    n = len(proof) // 3                                                                  # gpu_prove_once.py <c> 3
    agg = collections.OrderedDict()
    for d in proof[-n:]:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"] / 1000
        a[2] += d.get("sm__inst_executed_pipe_fma.sum", 0.0)
    tt, tf = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values()) or 1.0
    print("| kernel | launches | device time (us, serialised cold-cache) | share of time | FMA-pipe warp-instructions (M) | share of multiply-pipe work |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% | %.2f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tt, a[2] / 1e6, 100 * a[2] / tf))
    print("| **total** | %d | %.1f | | %.1f | |" % (n, tt, tf / 1e6))


FULL = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "registers"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (wide IMAD saturates at ~27)"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction")]


def full(path):
    rows = list(csv.reader(open(path)))
    H, U, data = rows[0], rows[1], rows[2:]
    kn, gs = H.index("Kernel Name"), H.index("Grid Size")
    cols = ["`%s` grid %s" % (re.sub(r"\(.*", "", r[kn]).replace("void ", ""), r[gs].replace(" ", "")) for r in data]
    print("| metric | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for name, label in FULL:
        if name not in H:
            continue
        i = H.index(name)
        vals = []
        for r in data:
            try:
                vals.append("%.2f %s" % (float(r[i].replace(",", "")), U[i] if U[i] not in ("%", "inst", "") else ""))
            except ValueError:
                vals.append(r[i])
        print("| %s | " % label + " | ".join(v.strip() for v in vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
