"""Turn the raw CSV of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
        --clock-control none --csv --log-file raw.csv python tools/layer_profile.py --warmup 0
into one row per kernel launch (profiles/rNN_ncu_forward_b1024.csv):  python tools/ncu_forward_csv.py raw.csv out.csv "comment" """
import csv, re, sys

raw, out, comment = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
lines = [l for l in open(raw) if not l.startswith("==")]
rows = {}
order = []
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in rows:
        rows[i] = {"kernel": re.sub(r"^void |eegldm::<unnamed>::|\(.*$", "", r["Kernel Name"]), "grid": r["Grid Size"]}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    u, n = r["Metric Unit"], r["Metric Name"]
    if n == "gpu__time_duration.sum":
        rows[i]["time_us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    elif n.startswith("dram__bytes_read"):
        rows[i]["rd"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    elif n.startswith("dram__bytes_write"):
        rows[i]["wr"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    elif n.startswith("gpu__dram_throughput"):
        rows[i]["dram_pct"] = v
    elif n.startswith("sm__pipe_tensor"):
        rows[i]["tensor_pct"] = v
    elif n.startswith("lts__t_bytes"):
        rows[i]["l2"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
with open(out, "w") as f:
    f.write("# " + comment + "\n")
    f.write("id,kernel,grid,time_us,dram_read_MB,dram_write_MB,dram_pct,tensor_pct,l2_bytes_MB\n")
    for i in order:
        r = rows[i]
        f.write(f'{i},{r["kernel"]},"{r["grid"]}",{r.get("time_us", 0):.1f},{r.get("rd", 0):.1f},{r.get("wr", 0):.1f},'
                f'{r.get("dram_pct", 0):.1f},{r.get("tensor_pct", 0):.1f},{r.get("l2", 0):.0f}\n')
