"""Summarise `ncu --page raw --csv` exports: one line per profiled launch with the metrics the rooflines quote."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; units = rows[1]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    tens = [i for i, h in enumerate(hdr) if "pipe_tensor" in h and "pct" in h]
    print("==", path)
    print("kernel,grid,block,regs,time_us,dram_read_MB,dram_write_MB,dram_pct,sm_pct,tensor_pct_max,l2_hit_pct,warps_active_pct")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:60]
        def g(k, scale=1.0):
            if k not in idx: return ""
            v = num(r[idx[k]]);
            if v is None: return ""
            u = units[idx[k]].lower()
            if k == "gpu__time_duration.sum":
                v = v / 1e3 if u.startswith("ns") else (v * 1e3 if u.startswith("ms") else (v * 1e6 if u == "s" or u == "second" else v))
            if "bytes" in k:
                v = v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
            return "%.2f" % (v * scale)
        tp = max([num(r[i]) or 0.0 for i in tens] + [0.0])
        print(",".join([name, g("launch__grid_size"), g("launch__block_size"), g("launch__registers_per_thread"), g("gpu__time_duration.sum"),
                        g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        g("sm__throughput.avg.pct_of_peak_sustained_elapsed"), "%.2f" % tp, g("lts__t_sector_hit_rate.pct"),
                        g("sm__warps_active.avg.pct_of_peak_sustained_active")]))
