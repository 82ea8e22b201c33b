"""Turns an `ncu --set full` capture of the rollout kernel (tools/ncu_rollout.sh) into the profiler-only numbers that
bench.py quotes (profiles/r02_ncu_constants.json): DRAM traffic per instance and pipe utilisations, with the capture's
source file, the number of instances it ran and the git head it was taken at.
    python tools/ncu_constants.py gpurun_out/r02_fast_full.ncu-rep --mode fast --instances 296 --git-head <sha>"""
import argparse, csv, io, json, os, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--mode", required=True)
ap.add_argument("--instances", type=int, required=True)
ap.add_argument("--nodes", type=int, default=200)
ap.add_argument("--group", type=int, default=200)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--decode-impl", default="")
ap.add_argument("--git-head", default="")
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                              "profiles", "r02_ncu_constants.json"))
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name, default=None):
    if name not in m:
        return default
    try:
        return float(m[name][0].replace(",", ""))
    except ValueError:
        return default


def to_bytes(name):
    v, u = num(name), m.get(name, ("", ""))[1].lower()
    if v is None:
        return None
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
entry = {
    "mode": a.mode, "nodes": a.nodes, "group": a.group, "layers": a.layers, "decode_impl": a.decode_impl,
    "instances": a.instances, "git_head": a.git_head, "source": os.path.basename(a.report),
    "kernel": m.get("Kernel Name", ("?", ""))[0],
    "duration_ms_under_ncu": (num("gpu__time_duration.sum") or 0) / 1e6 if m.get("gpu__time_duration.sum", ("", ""))[1] in ("nsecond", "ns") else num("gpu__time_duration.sum"),
    "dram_bytes_per_instance": None if rd is None or wr is None else (rd + wr) / a.instances,
    "pipes": {
        "issue_slots_pct": num("sm__inst_issued.avg.pct_of_peak_sustained_active") or num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "xu_mufu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "tensor_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        or num("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
        "registers_per_thread": num("launch__registers_per_thread"),
    },
}
table = {"captures": []}
if os.path.isfile(a.out):
    table = json.load(open(a.out))
table["captures"] = [e for e in table["captures"]
                     if (e["mode"], e["nodes"], e["group"], e["layers"], e.get("decode_impl", "")) !=
                     (a.mode, a.nodes, a.group, a.layers, a.decode_impl)] + [entry]
json.dump(table, open(a.out, "w"), indent=1)
print(json.dumps(entry, indent=1))
