#!/bin/bash
# Stages the reference's tests/ directory (read-only input, never committed: tests/_reference_tests is git-ignored) so
# that tests/test_reference_suite_gpu.py can run it on the GPU box, where /root/reference does not exist.
# usage: tools/stage_reference_tests.sh [stage|clean]
cd "$(dirname "$0")/.."
if [ "${1:-stage}" = "clean" ]; then rm -rf tests/_reference_tests; exit 0; fi
rm -rf tests/_reference_tests && mkdir -p tests/_reference_tests && cp -r /root/reference/tests/. tests/_reference_tests/
find tests/_reference_tests -name "__pycache__" -type d -prune -exec rm -rf {} +
