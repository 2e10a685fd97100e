#!/bin/bash
# Regenerates the raw counts behind profiles/sass_summary.txt (needs only the built library, no GPU).
cuobjdump -sass curvature_b200/libcurvature_b200.so > /tmp/crv_sass.txt
for m in UTCHMMA UTCBAR LDTM UTMALDG SYNCS HMMA; do printf "%-8s %s\n" $m "$(grep -c "\b$m" /tmp/crv_sass.txt)"; done
