#!/bin/bash
# last validation of the committed tree: full GPU suite + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02n_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1
tail -3 gpurun_out/r02n_pytest.log; tail -2 gpurun_out/r02n_smoke.log
