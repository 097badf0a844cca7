"""Shared pieces of the beam-search tests: the seeded logits generator, the oracle-side state of a run and the case list."""
import numpy as np

from oracle import beam_search_ref as BS

CASES = [
    dict(B=1, K=3, V=1000, Vp=1008, lens=[5]),                                       # input_demo.jsonl line 3: beam_width 3
    dict(B=2, K=4, V=2048, Vp=2048, lens=[6, 3], temperature=0.7),
    dict(B=3, K=2, V=5000, Vp=5056, lens=[4, 6, 2], repetition_penalty=1.3),
    dict(B=2, K=5, V=3000, Vp=3008, lens=[6, 6], diversity_rate=0.4, length_penalty=0.8, temperature=1.5, repetition_penalty=1.1),
    dict(B=1, K=8, V=100864, Vp=100864, lens=[4]),                                   # CodeFuse vocabulary, 99 parts
    dict(B=4, K=16, V=1500, Vp=1504, lens=[3, 5, 6, 2]),
]


def step_logits(B, K, Vp, end_id, step, seed):
    g = np.random.default_rng(seed * 1000003 + step)
    x = g.normal(0, 2.0, size=(B * K, Vp)).astype(np.float32)
    if step % 3 == 1:                       # end_id attractive for one row: beams finish at different times
        x[(step // 3) % (B * K), end_id] = 16.0
    x[:, 7] = x[:, 3]                       # exact ties: the smaller id must win (reduce_kernel_utils.cuh:275-322)
    return x


class OracleRun:
    """numpy state of one beam-search run, advanced by oracle/beam_search_ref.beam_step."""

    def __init__(self, B, K, V, Vp, max_in, out_len, lens, end_id, seed, **args):
        self.B, self.K, self.V, self.Vp, self.max_in, self.max_len, self.end_id, self.args = B, K, V, Vp, max_in, max_in + out_len, end_id, args
        BB = B * K
        g = np.random.default_rng(seed)
        self.ids0 = np.zeros((self.max_len, BB), dtype=np.int32)
        self.ids0[:max_in] = np.repeat(g.integers(0, V - 1, size=(B, max_in)), K, axis=0).T
        self.lens = np.repeat(np.asarray(lens, np.int32), K)
        fin, cum = BS.decoding_initialize(B, K)
        self.cum0 = cum.copy()
        self.ids, self.par = self.ids0.astype(np.int64), np.zeros((self.max_len, BB), np.int64)
        self.seq, self.fin, self.cum = np.full(BB, max_in - 1, np.int64), fin, cum
        self.ind = [np.zeros((BB, self.max_len), np.int64), np.zeros((BB, self.max_len), np.int64)]
        self.stop = args.get("stop_words")

    def logits(self, step, seed):
        return step_logits(self.B, self.K, self.Vp, self.end_id, step, seed)

    def advance(self, x, step):
        a = self.args
        par = (step - self.max_in) % 2
        BS.beam_step(x.copy(), step, self.ids, self.par, self.seq, self.fin, self.cum, self.ind[par], self.ind[1 - par], self.lens,
                     self.max_in, self.K, self.V, self.end_id, a.get("temperature", 1.0), a.get("repetition_penalty", 1.0),
                     a.get("diversity_rate", 0.0), a.get("length_penalty", 0.0), self.stop)
