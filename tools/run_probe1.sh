for f in 0/8 3/8 7/8; do
echo "== fake $f"
SHARD_FAKE=$f SHARD_STAGES=1 SHARD_MODES=block:8192,morton:4096,morton:1250000 python tools/shard_bench.py 320 31250 1 2>&1 | grep "rep 1: search"
done
