"""Turns the ncu CSVs written by profiles/ncu_capture.sh (gpurun_out/) into the tables of profiles/README.md.
Usage: python profiles/summarize.py [dir]   (default gpurun_out)"""
import collections
import csv
import re
import sys

D = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def rows(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    return [dict(zip(hdr, r)) for r in rd]


def launch_table(path, title):
    agg = collections.OrderedDict()
    for r in rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("fac::<unnamed>::", "").replace("void ", "")[:64]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}[r["Metric Unit"]]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("\n### %s (total %.2f ms)\n\n| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|" % (title, tot / 1e6))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
        print("| `%s` | %d | %.3f | %.1f | %.3f |" % (k, n, t / 1e6, t / n / 1e3, t / tot))


def raw_metrics(path, names):
    """`ncu --page raw --csv`: one row per launch, one column per metric, second line = units."""
    with open(path) as fh:
        rd = csv.reader(l for l in fh if l.startswith('"'))
        hdr, units = next(rd), next(rd)
        for row in rd:
            print("\nlaunch %s: %s  grid %s block %s" % (row[0], re.sub(r"\(.*", "", row[4])[:60], row[8], row[7]))
            for n in names:
                if n in hdr:
                    i = hdr.index(n)
                    print("  %-66s %s %s" % (n, row[i], units[i]))


METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__cluster_dim_x",
           "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]

if __name__ == "__main__":
    for name, title in (("launches_bf16x3.csv", "bench.py timed region, one step (WaveGlow.infer 8 x 10 s, bf16x3)"),
                        ("launches_tacotron.csv", "two Tacotron2.inference calls, 32 x 690 frames")):
        try:
            launch_table("%s/%s" % (D, name), title)
        except FileNotFoundError:
            pass
    for name in ("prof_bf16x3_raw.csv", "prof_decoder_raw.csv"):
        try:
            raw_metrics("%s/%s" % (D, name), METRICS)
        except FileNotFoundError:
            pass
