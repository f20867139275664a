"""profiles/roofline_traffic.json from the raw CSV of an `ncu --set full` capture of tools/profile_run.py:
DRAM bytes per tile and FP64-pipe utilisation of the four pipeline kernels.
usage: python tools/ncu_traffic.py raw.csv ntiles nb rank out.json "<source note>" """
import csv
import json
import sys

raw, T, nb, k, out, note = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]


def col(d, u, key):
    c = [h for h in hdr if h == key]
    if not c:
        return None
    v = float(d[c[0]].replace(",", "") or 0)
    unit = u[c[0]]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return v * mult


res = {"nb": nb, "rank": k, "tiles": T, "source": note}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = d["Kernel Name"].replace("void ", "").split("<")[0].split("(")[0].split("::")[-1]
    rd, wr = col(d, u, "dram__bytes_read.sum"), col(d, u, "dram__bytes_write.sum")
    ent = {"dram_bytes_per_tile": (rd + wr) / T, "dram_read": rd, "dram_write": wr,
           "duration_ns": col(d, u, "gpu__time_duration.sum"),
           "fp64_pipe_active_pct": col(d, u, "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"),
           "dmma_pct": col(d, u, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
           "dfma_pct": col(d, u, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_slots_busy_pct": col(d, u, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "l2_hit_pct": col(d, u, "lts__t_sector_hit_rate.pct")}
    res[name] = ent
res["tlr_stack_kernel"] = res.get("tlr_pg_kernel")
res["tiles_in_capture"] = T
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k2: (v["dram_bytes_per_tile"] if isinstance(v, dict) else v) for k2, v in res.items() if k2.startswith("tlr_")}))
