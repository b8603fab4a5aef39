"""profiles/umma_traffic.json from one `ncu --set full` capture of the surface kernel: DRAM bytes per launch, the library
digest the capture was taken with (bench.py refuses the value under any other build) and the engine's arithmetic.
  python tools/make_traffic.py gpurun_out/prof_x.ncu-rep  [kind 16|32]"""
import csv, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bogp_b200 import build

rep = sys.argv[1]
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(d[i].replace(",", "")) * scale for d in data]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
doc = {"kernel": data[0][hdr.index("Kernel Name")].split("(")[0], "launches_captured": len(data),
       "dram_bytes_read": sum(rd) / len(rd), "dram_bytes_write": sum(wr) / len(wr),
       "algorithmic_bytes": 4_000_000, "library_digest": build._digest(), "kind": kind,
       "note": "per launch; captured with --cache-control none inside the bench step (table kernel -> surface kernel), so "
               "the Bop tables the table kernel has just written are read from L2 as in an unprofiled run"}
(ROOT / "profiles" / "umma_traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
print(doc)
