EPA_B200_PIPE_DEBUG=1 python tools/files_bench.py 1000000 1 1 2>&1 | grep -v "^INFO" | tail -60
