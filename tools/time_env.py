"""python tools/time_env.py <env> <S> <A> <E> <T>: closed-loop (one CUDA graph) and fused timings of one config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sweep import run
r = run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
print(f"{r['env']} S={r['S']} A={r['A']} E={r['E']} T={r['T']}: closed {r['closed']:.3g}/s ({r['ms_closed']:.3f} ms)  fused {r['fused']:.3g}/s ({r['ms_fused']:.3f} ms)")
