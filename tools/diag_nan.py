import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import ChronoMindB200, default_config, last_counters
rng = np.random.default_rng(3)
data = datasets.unit_vectors(rng, 9000, 64)
q = datasets.unit_vectors(rng, 70, 64)
q[13, 5] = np.nan
wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, 10, mode="prepared")
for rep in range(4):
    st = ChronoMindB200.from_matrix(data, default_config(dimensions=64)).upload(0).set_exact_engine(2)
    ids, d = st.search_exact(q, 10)
    c = last_counters()
    gb, wb = d.view(np.uint32), wd.view(np.uint32)
    diff = np.argwhere(gb != wb)
    print("rep", rep, c, "ids equal", np.array_equal(ids, wi), "n diff", len(diff), diff[:12].tolist())
    for r, cidx in diff[:3]:
        print("   ", r, cidx, hex(gb[r, cidx]), hex(wb[r, cidx]), d[r, cidx], wd[r, cidx])
print("---- fp32 engine")
st = ChronoMindB200.from_matrix(data, default_config(dimensions=64)).upload(0).set_exact_engine(1)
ids, d = st.search_exact(q, 10)
print(ids[13], d[13].view(np.uint32))
ids, d = st.search_exact(q[13:14], 10)
print("single", ids[0], d[0].view(np.uint32))
import torch
from chronomind_b200.engine import native
import ctypes as C
qq = torch.from_numpy(q[13:14].copy()).cuda()
out = torch.empty((1, 64), dtype=torch.float32, device="cuda")
native().cm_metric_preprocess_dev(qq.data_ptr(), 1, 64, out.data_ptr(), 0, None)
torch.cuda.synchronize()
print("preprocessed NaN row bits:", out.cpu().numpy().view(np.uint32)[0, :8])
