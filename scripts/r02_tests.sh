#!/bin/bash
# round 2: the whole GPU suite (smoke gate first)
mkdir -p gpurun_out
T=${TAG:-r02o}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
timeout 1500 python -m pytest tests -q -m gpu --durations=8 ${PYA} 2>&1 | tail -40 | tee gpurun_out/pytest_${T}.log
