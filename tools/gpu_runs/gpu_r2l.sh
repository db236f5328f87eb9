#!/bin/bash
# round-2 final (1 GPU): smoke(), the full default bench line, compute-sanitizer on small cases of the new kernels
set -u
o=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $o/r2l_smoke.log 2>&1
timeout 1500 python bench.py > $o/r2l_bench_n1.json 2> $o/r2l_bench_n1.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu -x \
   "tests/test_gpu_dense.py::test_dense_kway_follows_the_oracle" "tests/test_gpu_snn.py::test_device_snn_equals_host_snn" \
   "tests/test_gpu_recursion.py::test_graph_split_equals_networkx_subgraphs" \
   "tests/test_gpu_recursion.py::test_concatenated_structured_models_anneal_like_separate_calls" > $o/r2l_memcheck.log 2>&1
echo "memcheck rc=$?" >> $o/r2l_memcheck.log
