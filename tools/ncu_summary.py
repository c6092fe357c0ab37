"""Summarise an .ncu-rep (raw metrics + source-level stall samples) on a box without a GPU.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__cycles_active.avg ", "sm__cycles_elapsed.avg ",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct",
        "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg ", "l1tex__m_xbar2l1tex_read_bytes.sum ",
        "sm__throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "smsp__inst_executed.sum ", "sm__pipe_tensor_cycles_active"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for kernel_row in rows[2:]:
        print("== kernel:", kernel_row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, kernel_row):
            if any(k.strip() in h for k in KEYS):
                print(f"  {h:75s} {v:>18s} {u}")
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    # the page starts with a 'Kernel Name' line, then the header
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], rows[hi + 1:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(r, k):
        try:
            return float(r[idx[k]])
        except (ValueError, IndexError):
            return 0.0
    total = sum(f(r, "# Samples") for r in data)
    print(f"== source: {len(data)} instructions, {total:.0f} samples")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
        s = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
        print(f"  {r[idx['Address']][-6:]} {f(r, '# Samples'):7.0f} {f(r, 'Instructions Executed'):10.0f}  "
              f"{r[idx['Source']][:78]:78s} {s[0][0]}={s[0][1]:.0f} {s[1][0]}={s[1][1]:.0f}")


if __name__ == "__main__":
    main()
