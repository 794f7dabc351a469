"""developer: the cfg5 TF32-vs-truncating-oracle report for several batch sizes (does the deviation grow with the chunks per CTA?)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_fullsize as T
import oracle.layers as Lm
orig = (Lm.conv2d, Lm.conv2d_wgrad, Lm.conv2d_dgrad)
for b in [int(a) for a in sys.argv[1:]] or [24, 48, 256]:
    Lm.conv2d, Lm.conv2d_wgrad, Lm.conv2d_dgrad = orig
    rep = T.cfg5_tf32_emulation_report(b, setattr)
    print(b, {k: rep[k] for k in ("loss", "13/weights", "10/gamma", "8/weights", "6/gamma", "4/weights", "0/weights")}, flush=True)
