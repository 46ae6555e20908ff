for share in 1 2 4 8; do for pool in 8 12; do echo "share $share"; PHOVO_POOL_SM_SHARE=$share python tools/bench_pool.py $pool 256 2>&1 | tail -1; done; done
