# e2e step time of bench.py against the upload chunking (RSC_CHUNK_PTS, RSC_FIRST_CHUNK_PTS)
for cfg in "2097152 2097152" "2097152 524288" "4194304 1048576"; do set -- $cfg; echo -n "chunk=$1 first=$2: "; RSC_CHUNK_PTS=$1 RSC_FIRST_CHUNK_PTS=$2 timeout 300 python bench.py --ransac none --no-cpu --no-cull 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done
