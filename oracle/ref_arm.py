"""Times the UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) on one
Yelp-shaped SML period (TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by sml_b200).

What runs is the reference's own code through its own public methods:
  ``meta_train.MF_train_onestage``        HOT LOOP A incl. its DataLoader(PreSampleDatast) (model/transfer.py:417-534)
  ``meta_train.transfer_train_onestage``  HOT LOOP B incl. SampleDaset + DataLoader          (:644-749)
  ``meta_train.updata``                   full-table transfer + load_MFbase_weight            (:884-902,945-959)
  ``transfer.test_model``                 candidate evaluation, DataLoader(testDataset, 1024) (evalution/evaluation2.py:8-26)
with ``--numworkers 0`` and the shims of SURVEY.md 8c (oracle/ref_harness.py).  A full CPU period
takes ~90-130 s, so each bench step runs a BOUNDED SAMPLE of every phase (a fraction of an MF epoch, of
a transfer epoch, of one evaluation, and one complete ``updata``) and composes one period from the
per-step / per-row times and the period's step counts; the measured fraction is reported.

    python -m oracle.ref_arm --device cpu|cuda [--scale S] [--seed N] [--rows R]    -> one JSON line
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

YELP = dict(n_users=59082, n_items=122816, rows=75000, n_neg=999)
HYPER = dict(multi_num=10, MF_epochs=1, TR_epochs=1, MF_batch_size=1024, TR_batch_size=256, MF_lr=0.01, l2=1e-6, TR_lr=0.001,
             TR_l2=1e-4, topK=20)


def period_counts(rows, h=HYPER):
    """(MF steps, transfer steps, full-table transfers, evaluations) of one train-only period."""
    mf_steps = -(-rows // h["MF_batch_size"]) * h["MF_epochs"] * h["multi_num"]
    tr_steps = -(-rows // h["TR_batch_size"]) * h["TR_epochs"] * h["multi_num"]
    updata = h["multi_num"] * (1 + h["TR_epochs"]) + 1
    evals = h["multi_num"] * (1 + h["MF_epochs"] + 1 + h["TR_epochs"])
    return mf_steps, tr_steps, updata, evals


class RefPeriod(object):
    """One reference ``meta_train`` on ``device`` plus synthetic period arrays sized for the sample."""

    def __init__(self, device="cpu", shape=YELP, seed=0, n_mf=12, n_tr=40, n_ev=3072):
        import torch
        from oracle import ref_harness
        self.torch = torch
        self.device = device
        self.ns = ns = ref_harness.load(device=device)
        U, I = shape["n_users"], shape["n_items"]
        self.U, self.I, self.rows = U, I, shape["rows"]
        parser = ns.main_yelp.get_parse()
        args = parser.parse_args([])
        args.data_name = "yelp"; args.numworkers = 0
        for k, v in HYPER.items():
            setattr(args, k, v)
        self.tmp = tempfile.mkdtemp(prefix="sml_ref_arm_")
        args.pre_model = os.path.join(self.tmp, "pre.pkl")
        torch.manual_seed(args.seed + seed); np.random.seed(args.seed + 2 + seed)
        pre = ns.MF.MFbasemode(U, I, 64)
        if device == "cuda":
            pre = pre.cuda()                        # "if your model not in cuda, try to put it into cuda" (model/transfer.py:320)
        torch.save(pre, args.pre_model)
        self.args = args
        with contextlib.redirect_stdout(io.StringIO()):
            self.meta = ns.transfer.meta_train(args, None, U, I, 64)
        rng = np.random.default_rng(seed)
        Bm, Bt = HYPER["MF_batch_size"], HYPER["TR_batch_size"]
        self.n_mf, self.n_tr, self.n_ev = n_mf, n_tr, n_ev
        n_t = n_mf * Bm
        # test-format rows (user, positive, 999 negatives) for the MF epoch sample and the evaluation sample
        self.set_t = np.concatenate([rng.integers(0, U, (n_t, 1)), rng.integers(0, I, (n_t, 1 + shape["n_neg"]))], 1).astype(np.int64)
        self.val = np.concatenate([rng.integers(0, U, (n_ev, 1)), rng.integers(0, I, (n_ev, 1 + shape["n_neg"]))], 1).astype(np.int64)
        # train-format rows (user, item) for the transfer epoch sample
        self.set_tt = np.stack([rng.integers(0, U, n_tr * Bt), rng.integers(0, I, n_tr * Bt)], 1).astype(np.int64)

    def sync(self):
        if self.device == "cuda":
            self.torch.cuda.synchronize()

    def _timed(self, fn):
        self.sync()
        t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            fn()
        self.sync()
        return time.perf_counter() - t

    def run(self):
        """One bounded sample of every phase -> per-unit times and the composed period."""
        m, a, ns, torch = self.meta, self.args, self.ns, self.torch
        m.save_MF_weight(save_as="last")
        t_mf = self._timed(lambda: m.MF_train_onestage(a, self.set_t, 0, val=None)) / self.n_mf
        m.MFbase.eval()
        m.save_MF_weight(save_as="hat")
        t_up = self._timed(m.updata)
        t_tr = self._timed(lambda: m.transfer_train_onestage(a, self.set_tt, 0, val=None)) / self.n_tr
        loader = torch.utils.data.DataLoader(ns.dataset2.testDataset(self.val), batch_size=1024, num_workers=0, pin_memory=False)
        t_ev_row = self._timed(lambda: ns.transfer.test_model(m.MFbase, loader, topK=a.topK)) / self.n_ev
        mf_steps, tr_steps, n_up, n_ev = period_counts(self.rows)
        parts = dict(mf=mf_steps * t_mf, tr=tr_steps * t_tr, updata=n_up * t_up, eval=n_ev * self.rows * t_ev_row)
        period = sum(parts.values())
        measured = self.n_mf * t_mf + self.n_tr * t_tr + t_up + self.n_ev * t_ev_row
        return dict(period_s=period, periods_per_s=1.0 / period, mf_step_ms=t_mf * 1e3, tr_step_ms=t_tr * 1e3, updata_ms=t_up * 1e3,
                    eval_rows_per_s=1.0 / t_ev_row, phase_s=parts, measured_s=measured, measured_fraction=measured / period)

    def sample_text(self):
        mf_steps, tr_steps, n_up, n_ev = period_counts(self.rows)
        return ("unmodified reference (%s): MF_train_onestage on %d rows = %d steps of %d (DataLoader, dense Adam on %dx64 + %dx64), "
                "transfer_train_onestage on %d rows = %d steps of %d, one full updata() (%d rows), test_model on %d rows x 1000; "
                "composed to one period = %d MF + %d TR steps + %d updata + %d evals of %d rows"
                % (self.device, self.n_mf * HYPER["MF_batch_size"], self.n_mf, HYPER["MF_batch_size"], self.U, self.I,
                   self.n_tr * HYPER["TR_batch_size"], self.n_tr, HYPER["TR_batch_size"], self.U + self.I, self.n_ev,
                   mf_steps, tr_steps, n_up, n_ev, self.rows))


def measure(device="cpu", scale=1.0, seed=0, rows=0, repeats=1, warm=True):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape = dict(YELP)
    if rows:
        shape["rows"] = rows
    gpu = device == "cuda"
    k = 8.0 if gpu else 1.0                                   # the GPU path is ~30x faster per step: sample more of it
    p = RefPeriod(device, shape, seed, n_mf=max(2, int(12 * scale * k)), n_tr=max(4, int(40 * scale * k)),
                  n_ev=max(1024, int(3072 * scale * k)))
    if warm:                                                  # allocator / thread pool / cuDNN algorithm choice
        w = RefPeriod.__new__(RefPeriod)
        w.__dict__.update(p.__dict__)
        w.n_mf, w.n_tr, w.n_ev = 1, 2, 1024
        w.set_t, w.set_tt, w.val = p.set_t[:1024], p.set_tt[:512], p.val[:1024]
        w.run()
    outs = [p.run() for _ in range(repeats)]
    best = min(outs, key=lambda o: o["period_s"])
    best.update(cores=cores, device=device, sample=p.sample_text(), kind="reference")
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(measure(a.device, a.scale, a.seed, a.rows, a.repeats)))


if __name__ == "__main__":
    main()
