#!/bin/bash
# rebuild libsifnn_b200.so in-tree (sm_100a); exits non-zero on failure
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys
sys.path.insert(0, '.')
import sifnn_b200
print(sifnn_b200.build(force=True))
lib = sifnn_b200.load()
print("symbols ok, version", lib.sifnn_version())
PY
