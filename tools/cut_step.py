"""Keeps the LAST `per_step` kernel launches of an ncu launch list: the timed step of
`bench.py --steps 1 --warmup 1 --only-step` (setup kernels and the warm-up step come before it)."""
import sys

lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
per = int(sys.argv[2])
sys.stdout.write(lines[0])
sys.stdout.writelines(lines[1:][-per:])
