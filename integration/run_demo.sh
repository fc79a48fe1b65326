#!/bin/bash
# Drop-in demonstration on a GPU box: the reference's own drivers (svd_feature.cpp,
# svd_feature_infer.cpp, apex_svd_data.cpp -- unmodified) linked against the GPU trainer run
# demo/basicMF/run.sh on the committed 4-row example and must print the reference CLI's
# predictions (tests/golden/ref_cli_pred.txt).  Usage: integration/run_demo.sh [workdir]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
WORK="${1:-$(mktemp -d)}"
mkdir -p "$WORK" && cd "$WORK"
cp "$ROOT/tests/golden/ua.base.buffer" ua.base.buffer
python - "$ROOT" <<'PY'
import sys
sys.path.insert(0, sys.argv[1])
from svdfeature_b200 import buffer_io
buffer_io.write_feature_buffer("ua.test.buffer", buffer_io.parse_feature_text(sys.argv[1] + "/tests/golden/ua.test.example.txt"))
PY
cat > basicMF.conf <<'CONF'
# demo/basicMF/basicMF.conf of the reference + one GPU key
base_score = 3
learning_rate = 0.005
wd_item       = 0.004
wd_user       = 0.004
num_item   = 1682
num_user   = 943
num_global = 0
num_factor = 64
active_type = 0
test:buffer_feature="ua.test.buffer"
buffer_feature = "ua.base.buffer"
model_out_folder="./"
gpu:mode = exact
CONF
"$HERE/_build/svd_feature_gpu" basicMF.conf num_round=40 > train.log 2>&1
"$HERE/_build/svd_feature_infer_gpu" basicMF.conf pred=40 > infer.log 2>&1
echo "--- pred.txt (GPU trainer behind the reference CLI)"; cat pred.txt
echo "--- tests/golden/ref_cli_pred.txt (reference CLI)"; cat "$ROOT/tests/golden/ref_cli_pred.txt"
cmp pred.txt "$ROOT/tests/golden/ref_cli_pred.txt" && echo "IDENTICAL predictions"
sha256sum 0040.model | cut -d' ' -f1; python -c "import json;print(json.load(open('$ROOT/tests/golden/ref_cli_0040.model.json'))['sha256'])"
