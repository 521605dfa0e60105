"""Write profiles/scramble3_dram_bytes_per_launch.json from an `ncu --set full` capture of K1p
(tools/run_kernels.py scramble3), stamped with the fingerprint of the kernel's sources so that bench.py
only quotes it for the build it was taken from:   python tools/update_traffic.py gpurun_out/k1p.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    row = [r for r in data if "scramble_pairs_kernel<3, 30, 2, 1>" in r[col["Kernel Name"]]][-1]

    def nbytes(key):
        v, u = float(row[col[key]]), units[col[key]].lower()
        return int(round(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]))

    rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    rec = {"kernel": "scramble_pairs_kernel<3,30,2,plain>", "instances_per_launch": 8 * 2 ** 20, "depth": 30,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": 8 * 2 ** 20 * 89,
           "from": "ncu --set full --clock-control none on `python tools/run_kernels.py scramble3 --iters 3` (%s)" % os.path.basename(rep),
           "source_sha16": bench.source_sha16()}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "scramble3_dram_bytes_per_launch.json"), "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
