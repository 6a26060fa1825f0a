"""Logit error of the encrypted ResNet-20 against the float64 model (oracle/plain_model.py) for a few images, in the
current key-switching mode ($B200CKKS_HYBRID_KS).  python tools/resnet_accuracy.py [n_images]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import plain_model as pm
from b200ckks import synthetic
from b200ckks.app import App

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bits = [51] + [46] * 16 + [51] * 14 + [51]
s = App().session(16, bits, hamming_weight=192)
w = synthetic.random_weights(20, seed=0)
net = s.resnet(20, w)
imgs = np.stack([synthetic.synthetic_image(i) for i in range(n)])
got = net.infer_batch(imgs, min(n, 4))
for i in range(n):
    want = pm.resnet_forward(20, w, imgs[i])
    print(f"image {i}: max |logit - model| = {np.abs(got[i] - want).max():.2e}  argmax {int(np.argmax(got[i]))} / {int(np.argmax(want))}")
print("key residency GiB", s.key_residency()[0] / 2 ** 30)
s.close()
