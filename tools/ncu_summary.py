"""Summarise `ncu --set full ... --page raw --csv` output (one row per profiled launch) into the text committed under profiles/
and refresh profiles/ncu_traffic.json (dram read + write per launch of the kernels bench.py's rooflines name).

    ncu -i x.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [--traffic profiles/ncu_traffic.json]
"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def get(r, k, default=""):
    return r[idx[k]] if k in idx else default


def f(r, k):
    try:
        return float(get(r, k).replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(r, k):
    v, u = f(r, k), units[idx[k]] if k in idx else ""
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(r, k):
    v, u = f(r, k), units[idx[k]] if k in idx else ""
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
groups = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", get(r, "Kernel Name").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))
    groups.setdefault((name, get(r, "Grid Size"), get(r, "Block Size")), []).append(r)
traffic = {}
for (name, grid, block), rs in groups.items():
    r = rs[-1]                                   # the last profiled launch of the group
    st = sorted(((f(r, k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stall_keys), reverse=True)
    tot = sum(v for v, _ in st if v == v) or 1.0
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    print(f"=== {name}   grid {grid} block {block}   ({len(rs)} launch(es) profiled)")
    print(f"    duration                         {to_us(r, 'gpu__time_duration.sum'):10.2f} us")
    print(f"    dram read / write                {rd / 1e6:10.3f} / {wr / 1e6:.3f} MB   ({(rd + wr) / 1e3 / max(to_us(r, 'gpu__time_duration.sum'), 1e-9):.0f} GB/s)")
    print(f"    tensor pipe active               {f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):10.2f} % of active cycles")
    print(f"    warps active                     {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):10.2f} %")
    print(f"    issue slots busy                 {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):10.2f} %")
    print(f"    regs/thread, dyn smem/CTA        {get(r, 'launch__registers_per_thread')}, {get(r, 'launch__shared_mem_per_block_dynamic')} {units[idx['launch__shared_mem_per_block_dynamic']] if 'launch__shared_mem_per_block_dynamic' in idx else ''}")
    print(f"    cluster size                     {get(r, 'launch__cluster_size', '1')}")
    print(f"    L2 hit / L1 hit                  {f(r, 'lts__t_sector_hit_rate.pct'):.1f} % / {f(r, 'l1tex__t_sector_hit_rate.pct'):.1f} %")
    print("    top stall reasons                " + ", ".join(f"{k} {100 * v / tot:.0f}%" for v, k in st[:6]))
    print()
    traffic[name] = rd + wr
for a in sys.argv[2:]:
    pass
if "--traffic" in sys.argv:
    out = sys.argv[sys.argv.index("--traffic") + 1]
    m = {"decoder_step_projection": "dec_proj_kernel", "ctc_prefix_full_vocab": "ctc_prefix_full_kernel", "decode_source_attention": "dec_attn_stream_kernel<1, 4>",
         "encoder_ffn1_gemm_bf16_tcgen05": "gemm_tc_kernel<256>"}
    try:
        cur = json.load(open(out))
    except (OSError, ValueError):
        cur = {}
    cur["_source"] = f"tools/ncu_summary.py {path}: dram__bytes_read.sum + dram__bytes_write.sum of the last profiled launch of each kernel (ncu --set full)"
    for k, kn in m.items():
        hit = [v for n, v in traffic.items() if n.startswith(kn)]
        if hit:
            cur[k] = hit[-1]
    json.dump(cur, open(out, "w"), indent=1)
