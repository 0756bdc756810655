#!/bin/bash
# Round-end check on the GPU box: the whole GPU suite, smoke(), then the profile set (tools/profile_r2.sh).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/profile_r2.sh
