for v in 0 150 300 450; do echo "PF=$v"; INSTA_HASH_PF=$v timeout 200 python tools/time_hash.py 2>&1 | tail -1; done
