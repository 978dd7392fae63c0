#!/bin/bash
# round-2 GPU call 11: kNN after the epilogue fix + second pass, full-size k-means E-step test, tightened gradient test
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py "tests/test_kmeans_gpu.py::test_assign_full_baseline_size_sampled_rows_vs_oracle" "tests/test_kmeans_gpu.py::test_run_kmeans_save_load_and_decode_json" \
   "tests/test_model_gpu.py::test_training_gradients_match_oracle" -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02j_tests.log; tail -25 gpurun_out/r02j_tests.log | cut -c1-260
timeout 600 python tools/knn_quick.py 2>&1 | tail -8
timeout 900 python bench.py --workload knn > gpurun_out/r02j_bench_knn.json 2> gpurun_out/r02j_bench_knn.err; tail -c 1500 gpurun_out/r02j_bench_knn.json
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "roi or pooler or tap" 2>&1 | tail -5
