#!/bin/bash
# SASS of the default substep kernel instance (k_substep_seg<7,1,14,3,3,1>) plus an opcode histogram,
# from the object file the library is linked from.  Usage: bash tools/dump_sass.sh [tag]
set -e
cd "$(dirname "$0")/.."
tag=${1:-r02}
obj=eutropia_b200/csrc/_build/stencil_seg.o
fun=_ZN3eut13k_substep_segILi7ELi1ELi14ELi3ELi3ELi1EEEvPdllPKiPK4int2PK5uint4S9_PKdS3_S3_iiiiiiS3_
out=profiles/${tag}_k_substep_seg_7_1_14_3_3_1.sass
cuobjdump -sass -fun $fun $obj | grep -E "^\s+/\*[0-9a-f]{4}\*/|Function" | sed -E 's@/\* 0x[0-9a-f]+ \*/@@; s/[[:space:]]+$//' > $out
{
  echo "# opcode histogram of $out"
  grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" $out | awk '{print $NF}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn
} > profiles/${tag}_k_substep_seg_opcodes.txt
wc -l $out; head -30 profiles/${tag}_k_substep_seg_opcodes.txt
