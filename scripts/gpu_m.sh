#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
FAABRIC_TEST_WATCHDOG_SECS=40 timeout 120 build/bin/faabric_tests --tag gpu > gpurun_out/m_cpp_gpu.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|====" gpurun_out/m_cpp_gpu.log | tail -15; grep -A8 "FAIL" gpurun_out/m_cpp_gpu.log | head -30
