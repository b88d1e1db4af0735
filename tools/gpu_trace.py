"""Device-side timeline of one whole train step (cfg2): %globaltimer stamps at the fork / join points of the step's streams
(debug bit 13, csrc/api.cu k_stamp) + how many chunks the background encoder-Adam streamer claimed.  GPU box only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from spotify_recsys_challenge_2018_b200.models.DAEs import DAE
from tests.gpu_util import model_buf

NAMES = ["start(A)", "encode_end", "ybits_end", "decode_end", "dh_end", "dec_begin", "dec_end", "bg_begin", "bg_end", "tail_end",
         "rest_begin", "rest_end", "g1_first_cta", "g1_last_cta", "step_end", "prepared"]


def main():
    flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    T, A, H, B, tied = bench.WORKLOADS["cfg2"]

    class C:
        pass
    c = C()
    c.save = "/tmp/w"; c.batch = B; c.n_input = T + A; c.n_tracks = T; c.hidden = H; c.lr = bench.LR; c.reg_lambda = 0.0
    c.initval = "NULL"; c.seed = 0
    m = DAE(c).fit()
    m.set_debug(8192 | flags)
    batches = bench.make_batches("cfg2", 4, seed=7)
    m.stage_batch(0, *batches[0]); m.stage_batch(1, *batches[1])
    rows = []
    for i in range(30):
        m.train_step_staged(i & 1, bench.KP, bench.KP_IN)
        m.restage(i & 1)
        if i >= 10:
            m.sync_cost()
            t = model_buf(m, "trace", torch.int64).cpu().numpy()[:len(NAMES)].astype(np.float64)
            ctl = int(model_buf(m, "bg_ctl", torch.int32)[0].item()) & 0x3fffffff
            rows.append(np.concatenate([(t - t[0]) / 1e3, [ctl]]))
    r = np.median(np.array(rows), axis=0)
    print("flags", flags, "median over", len(rows), "steps (us after barrier A):")
    for n, v in zip(NAMES, r):
        print("  %-12s %9.1f" % (n, v))
    n_chunks = (m.n_input + 127) // 128 * 128 // 32
    print("  bg chunks claimed %d of %d (%.1f %% of the encoder rows)" % (r[-1], n_chunks, 100.0 * r[-1] / n_chunks))
    m.close()


if __name__ == "__main__":
    main()
