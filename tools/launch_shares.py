"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family and print launches, total time and share.
usage: python tools/launch_shares.py <launches.csv> [steps]      (steps: graph replays inside the profiled range, for the header)"""
import collections
import csv
import sys

FAMILIES = [
    ("cuDNN / cuBLAS (convolutions, fc)", ("cutlass", "xmma", "cudnn", "gemm_bf16", "sm80_", "sm90_", "sm100_", "nvjet", "gemv", "cublas")),
    ("msml bn_bwd", ("bn_bwd",)),
    ("msml bn_fwd", ("bn_fwd", "bn_apply", "bn_eval")),
    ("msml sgd_flat (backbone optimizer)", ("sgd_flat",)),
    ("msml pfc_sgd_update", ("pfc_sgd",)),
    ("msml fm_gate", ("fm_gate",)),
    ("msml fm_cat", ("fm_cat",)),
    ("msml head GEMMs (tcgen05)", ("msml::gemm", "gemm_kernel", "gemm_pair_kernel")),
    ("msml accum_bf16_multi", ("accum_bf16",)),
    ("msml dap", ("dap_",)),
    ("msml head small kernels", ("msml::",)),
    ("NCCL", ("nccl",)),
]


def family(name):
    for fam, keys in FAMILIES:
        if any(k in name for k in keys):
            return fam
    return "ATen glue (elementwise, copies, reductions, fills)"


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or r is hdr:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}.get(d["Metric Unit"], 1.0)
        a = agg.setdefault(family(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("# ncu launch list of the timed region (%d graph replays), grouped by kernel family; source: %s" % (steps, path))
    print("# cold-cache serialised durations: compare SHARES, not absolutes")
    print("# family, launches (%d steps), total_us (%d steps), share" % (steps, steps))
    for fam, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %5d %10.1f %6.1f%%" % (fam, n, t, 100.0 * t / total))
    print("%-58s %5d %10.1f" % ("total", sum(a[0] for a in agg.values()), total))


if __name__ == "__main__":
    main()
