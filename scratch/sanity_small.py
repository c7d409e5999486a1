import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle
from common import make_mask, synth_qkv, assert_close
from gpu_common import pkg, to_dev, run_both
P = pkg()
# one small case per kernel family
Q, K, V = synth_qkv(128, 1, 777, 8, 2, n_batch=3); run_both(Q, K, V, make_mask("noise", 1, 777)); print(P.last_dispatch())
Q, K, V = synth_qkv(128, 1, 776, 8, 2, n_batch=3); run_both(Q, K, V, None, q8=True); print(P.last_dispatch())
Q, K, V = synth_qkv(128, 40, 300, 8, 2); run_both(Q, K, V, make_mask("causal", 40, 300), flags=P.FLAG_CAUSAL); print(P.last_dispatch())
Q, K, V = synth_qkv(128, 300, 500, 6, 2, n_batch=2); run_both(Q, K, V, make_mask("causal", 300, 500), flags=P.FLAG_CAUSAL); print(P.last_dispatch())
Q, K, V = synth_qkv(128, 200, 257, 4, 2); run_both(Q, K, V, make_mask("noise", 200, 257)); print(P.last_dispatch())
print("sanity ok")
