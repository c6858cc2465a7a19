#!/bin/bash
# Developer helper: build a kernel variant of libsrt.so (production-mode instantiations of the default spectral
# width only) for A/B timing.
#   scripts/build_variant.sh NAME [-DFOO=1 ...]   ->  build/variants/libsrt_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
python - "$name" "$@" <<'PY'
import sys
sys.path.insert(0, ".")
import __graft_entry__ as g
name, defines = sys.argv[1], sys.argv[2:]
g.build_libsrt(f"build/variants/libsrt_{name}.so", defines=["-DSRT_DEV_MINIMAL", "-DSRT_DEV_ONLY_NL8", *defines],
               units=["srt_api.cu", "srt_resident_nl8.cu"], force=True, obj_tag="variant_" + name)
PY
