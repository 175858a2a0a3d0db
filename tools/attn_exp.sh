for d in 0 1 2 4 8 16 32 63; do echo "== dbg=$d"; DHG_OPTS=attn_dbg=$d python tools/attn_bench.py 1024 2>&1 | grep -E "self L1|self L3|cross L1" | sed 's/impl0.*impl1/impl1/'; done
