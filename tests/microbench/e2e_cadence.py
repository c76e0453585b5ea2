"""Inter-yield times of RanMtgEncDecDataset.host_tensor_batches (diagnostic for the e2e number)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mtgvision_b200 import synth
from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
from mtgvision_b200.encoder_train import RanMtgEncDecDataset
P = 512
cards = synth.make_card_pool(1024, workers=16); bgs = synth.make_bg_pool(1024, workers=16)
ds = RanMtgEncDecDataset(P, paired=True, targets=False, mtg=SyntheticBgFgMtgImages(pool=cards), ilsvrc=IlsvrcImages(images=bgs),
                         device=0, out_dtype="float16", seed=1)
hc = torch.from_numpy(cards.images[:P]).pin_memory(); hb = torch.from_numpy(np.stack(bgs[:P])).pin_memory()
def feed(k):
    for _ in range(k): yield hc, hb
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    for res in ds.host_tensor_batches(feed(12)):
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("rep", rep, "total %.1f ms" % (tot * 1e3), "yields at", [round(t * 1e3, 1) for t in ts])
