"""Build profiles/r02_ncu_traffic.json (bench.py's roofline.traffic source) from the raw pages of
`ncu --set full` captures of one eager C2 step per compute mode:

    ncu -i c2_bf16.ncu-rep --page raw --csv > c2_bf16_raw.csv
    python tools/ncu_traffic.py bf16=c2_bf16_raw.csv tf32=c2_tf32_raw.csv > profiles/r02_ncu_traffic.json

Per op group of bench.py (its `kernel_time_shares` keys at the C2 shapes) the value is
dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), averaged over the launches of that
kernel in the capture; the file records the library digest the capture was taken on."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel-name pattern -> bench.py op group (C2 shapes: M = 8000)
GROUPS = [
    (r"ffn_fused_kernel", "ffn_fused[(8000, 256), (2048, 256)]"),
    (r"relpos_attn_tc_kernel", "relpos_attn[(8000, 768), (499, 256)]"),
    (r"csgu_(conv|stats|onepass)_kernel", "csgu[(8000, 2048), (1024,)]"),     # summed over the pair
    (r"merge_scores_kernel", "merge_scores[(8000, 256), (8000, 1024)]"),
    (r"ctc_loss_kernel", "ctc_loss[(32, 250, 41), (32, 100)]"),
    (r"ctc_head_kernel", "ctc_head[(8000, 256), (41, 256)]"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = collections.defaultdict(lambda: [0, 0.0])
    pair = collections.defaultdict(int)
    for r in data:
        name = r[col["Kernel Name"]]
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0)
        for pat, key in GROUPS:
            if re.search(pat, name):
                out[key][1] += rd + wr
                if "csgu" in pat:
                    pair[re.search(pat, name).group(1)] += 1
                    out[key][0] = max(pair.values())          # launches of the op = launches per kernel
                else:
                    out[key][0] += 1
                break
    return {k: v[1] / max(1, v[0]) for k, v in out.items()}


def main():
    digest = None
    try:
        digest = open(os.path.join(ROOT, "tailored_avsr_b200", "libtavsr_sm100.so.digest")).read().strip()[:16]
    except OSError:
        pass
    sys.path.insert(0, ROOT)
    import bench
    res = {"lib_digest": digest, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "how": "ncu --set full --clock-control none on one eager C2 step per mode (tools/final2.sh)",
           "kernel_source_digests": {g: bench.kernel_source_digest(g) for g in bench.KERNEL_SOURCES},
           "kernel_source_files": bench.KERNEL_SOURCES}
    for arg in sys.argv[1:]:
        mode, path = arg.split("=", 1)
        res[mode] = load(path)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
