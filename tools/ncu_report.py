"""Summarise an `ncu --set full --import-source on` report without the GUI: headline metrics per kernel, the warp-stall
breakdown of the whole kernel and its hottest SASS instructions (needs -lineinfo for the source page).

    python tools/ncu_report.py gpurun_out/r01b_voc_c128k3.ncu-rep [--top 20] [--json out.json]
"""
import argparse
import csv
import io
import json
import subprocess
import sys

HEAD = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def ncu_csv(report, page):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        sys.exit(f"ncu failed on {report}: {out.stderr[:300]}")
    return list(csv.reader(io.StringIO(out.stdout)))


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return None if scale is None else float(v) * scale


def to_seconds(v, unit):
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}.get(unit)
    return None if scale is None else float(v) * scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--json")
    args = ap.parse_args()

    raw = ncu_csv(args.report, "raw")
    hdr, units, kernels = raw[0], raw[1], raw[2:]
    result = []
    for vals in kernels:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        rec = {"kernel": name}
        print("=" * 100)
        print(name[:140])
        for key, label in HEAD:
            if d.get(key, "") != "":
                rec[key] = {"value": d[key], "unit": u.get(key, "")}
                print(f"  {label:24s} {d[key]} {u.get(key, '')}")
        t = to_seconds(d.get("gpu__time_duration.sum", "nan"), u.get("gpu__time_duration.sum", ""))
        br = to_bytes(d.get("dram__bytes_read.sum", "nan"), u.get("dram__bytes_read.sum", ""))
        bw = to_bytes(d.get("dram__bytes_write.sum", "nan"), u.get("dram__bytes_write.sum", ""))
        if t and br is not None and bw is not None:
            rec["dram_tb_per_s"] = (br + bw) / t / 1e12
            print(f"  {'DRAM bandwidth':24s} {(br + bw) / t / 1e12:.2f} TB/s  ({(br + bw) / 1e6:.1f} MB per launch)")
        result.append(rec)

    # source page: one table per captured kernel is concatenated; summarise the first (reports here hold 1-2 kernels)
    src = ncu_csv(args.report, "source")
    try:
        h = next(i for i, r in enumerate(src) if "# Samples" in r)
    except StopIteration:
        print("(no source page: capture with --import-source on and build with -lineinfo)")
        h = None
    if h is not None:
        hdr = src[h]
        i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
        rows = []
        for r in src[h + 1:]:
            if len(r) != len(hdr):
                continue
            if not r[i_s].isdigit():            # header of the next kernel's table: keep the first kernel only
                break
            rows.append(r)
        stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        body = [r for r in rows if "EXIT" not in r[i_src]]          # idle warps park at the final barrier / EXIT
        tot = {}
        for r in body:
            for c in stall:
                tot[hdr[c]] = tot.get(hdr[c], 0) + int(r[c] or 0)
        n = sum(tot.values()) or 1
        print("-" * 100)
        print("warp stall samples (all warps, EXIT excluded):",
              ", ".join(f"{k[6:]} {100 * v // n}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        top = sorted(range(len(body)), key=lambda i: -int(body[i][i_s] or 0))[:args.top]
        print(f"hottest {len(top)} instructions (index in SASS order, samples, instruction, top stall reasons):")
        hot = []
        for i in sorted(top):
            r = body[i]
            st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall), reverse=True)[:2]
            hot.append({"index": i, "samples": int(r[i_s] or 0), "sass": r[i_src].strip(), "stalls": st})
            print(f"  {i:5d} {r[i_s]:>7s}  {r[i_src].strip()[:70]:70s} {st}")
        result.append({"stall_breakdown": tot, "hot": hot, "sass_instructions": len(rows)})
    if args.json:
        with open(args.json, "w") as f:
            json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
