#!/bin/bash
# Instruction mix of one kernel of libdark_bwt.so:  tools/sass_mix.sh <mangled-name-substring>
cuobjdump -sass dark_b200/lib/libdark_bwt.so | awk -v pat="$1" '/Function :/{p = index($0, pat) > 0} p' > /tmp/kern.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/kern.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*//' | awk '{ if ($1 ~ /^@/) print $2; else print $1}' | sed 's/[.;].*//' | sort | uniq -c | sort -rn | head -${2:-45}
echo "total: $(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/kern.sass)"
