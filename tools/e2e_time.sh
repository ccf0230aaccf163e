python - <<'PY'
import sys; sys.path.insert(0,'.')
from pathlib import Path
from kamino_b200 import synth
sp=synth.bacterial(100); synth.write_fasta_dir(sp, Path('/tmp/e2e_in'))
PY
mkdir -p /tmp/e2e_out
for i in 1 2; do KAMINO_TIMING=1 ./kamino_b200/_lib/kamino-b200 -i /tmp/e2e_in -t 16 --nj -o /tmp/e2e_out/k 2>&1 | grep time; echo; done
