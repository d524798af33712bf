"""Run the UNMODIFIED reference entry point (model/Run.py) with the B200-native module standing in for
model/Pretrain_model/GPTST.py -- the drop-in of SURVEY.md section 8b exercised end to end.

    python tools/run_reference.py -dataset PEMS08 -mode pretrain -batch_size 4 -epochs 1          # BASELINE.json configs[0]
    python tools/run_reference.py -dataset PEMS08 -mode eval -model STGCN --epochs 1              # config 5 (frozen encoder)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_reference.py -dataset PEMS08 -mode pretrain ...

Nothing of the reference is edited.  The launcher only
  * pre-seeds both import names of the model file (`model.Pretrain_model.GPTST` for Run.py:11 and `Pretrain_model.GPTST` for
    model/Model.py:3) with `gptst_b200.GPTST`, so Run.py, BasicTrainer.py, lib/ and conf/ run as they are;
  * restores `numpy.mat` (removed in NumPy 2; model/STGCN/args.py:25-43 uses it) -- an environment shim, SURVEY.md section 3.3;
  * under torchrun: `torch.cuda.set_device(LOCAL_RANK)` before anything touches CUDA (Run.py:27's torch.device("cuda") and
    lib/dataloader.py:94's torch.cuda.FloatTensor then land on the local GPU), every rank keeps the same seed so the loaders
    shuffle identically, each batch is sliced `[rank::world]` by wrapping `BasicTrainer.Trainer`'s loaders, the parameters
    are broadcast from rank 0 after Run.py's init loop and ONE flat gradient all-reduce runs at the end of every backward
    (`gptst_b200.dp.FlatGradAllReduce.attach`), i.e. before BasicTrainer.py:96's clip_grad_norm_.

`--config5-parity` (BASELINE.json configs[4]): runs `-mode eval -model STGCN ...` with the reference's own
encoder, so ONE downstream head is trained once; after training the frozen encoder inside that same `Enhance_model` is swapped
for the B200-native module (same state_dict) and the test-set MAE is computed through Model.py:106-117 with each encoder:
`[config5] test MAE reference-encoder .. native-encoder .. |dMAE| ..`  (north_star: within 1e-3).

`--reference-module` runs the reference's OWN GPTST.py instead (A/B on the same box: same command line, same log lines).
The reference root is `$GPTST_REFERENCE_ROOT`, `<repo>/baseline/_ref` (tools/install_reference.sh) or /root/reference; it must be
writable (Run.py writes model/SAVE/<dataset>/).
"""
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_root():
    for r in (os.environ.get("GPTST_REFERENCE_ROOT"), os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if r and os.path.isfile(os.path.join(r, "model", "Run.py")):
            return r
    raise SystemExit("run_reference.py: no reference tree found (run tools/install_reference.sh, or set GPTST_REFERENCE_ROOT)")


class _RankSlicedLoader:
    """The reference's DataLoader with every batch cut to this rank's rows (same length, same order on every rank)."""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            yield tuple(t[self.rank::self.world] for t in batch)

    def __getattr__(self, name):
        return getattr(self.loader, name)


def main():
    argv = sys.argv[1:]
    config5 = "--config5-parity" in argv
    use_reference_module = "--reference-module" in argv or config5
    argv = [a for a in argv if a not in ("--reference-module", "--config5-parity")]
    root = find_root()
    import numpy
    if not hasattr(numpy, "mat"):
        numpy.mat = numpy.asmatrix
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if not use_reference_module:
        sys.path.insert(0, REPO)
        import gptst_b200  # noqa: F401  (import shim for the package directory gpt-st_b200/)
        from gptst_b200 import GPTST as native
        sys.modules["model.Pretrain_model.GPTST"] = native
        sys.modules["Pretrain_model.GPTST"] = native
        print(f"[run_reference] GPTST_Model <- {native.__file__}", flush=True)
    model_dir = os.path.join(root, "model")
    os.chdir(model_dir)                       # the reference resolves ../conf and ../data relative to model/
    sys.path.insert(0, model_dir)             # what `python Run.py` would put there
    if world > 1:
        from gptst_b200 import dp
        dp.init_from_env()
        sys.path.append(root)
        import model.BasicTrainer as BT       # the reference's trainer, unmodified; wrap its constructor only
        orig_init = BT.Trainer.__init__

        def init(self, model, loss, loss_kl, optimizer, train_loader, val_loader, test_loader, *a, **kw):
            dp.broadcast_parameters(model)
            self._gptst_reducer = dp.FlatGradAllReduce(model.parameters()).attach()
            wrap = lambda ld: None if ld is None else _RankSlicedLoader(ld, rank, world)
            orig_init(self, model, loss, loss_kl, optimizer, wrap(train_loader), wrap(val_loader), wrap(test_loader), *a, **kw)

        BT.Trainer.__init__ = init
        print(f"[run_reference] rank {rank}/{world} on cuda:{torch.cuda.current_device()}", flush=True)
    if config5:
        sys.path.append(root)
        sys.path.insert(0, REPO)
        import model.BasicTrainer as BT
        orig_train = BT.Trainer.train

        def train(self):
            orig_train(self)
            import gptst_b200  # noqa: F401
            from gptst_b200.GPTST import GPTST_Model as Native
            from lib.metrics import All_Metrics
            net, args = self.model, self.args
            ref_enc = net.pretrain_model

            def predict():
                net.eval()
                ys, ts = [], []
                with torch.no_grad():
                    for data, target in self.test_loader:
                        data = data[..., :args.input_base_dim + args.input_extra_dim]
                        ys.append(net(data, label=None)[0])
                        ts.append(target[..., :args.output_dim])
                return self.scaler.inverse_transform(torch.cat(ys, 0)), self.scaler.inverse_transform(torch.cat(ts, 0))

            yp_ref, yt = predict()
            nat = Native(ref_enc_args(ref_enc, args)).to(yp_ref.device)
            nat.load_state_dict(ref_enc.state_dict(), strict=True)
            for p in nat.parameters():
                p.requires_grad = False
            net.pretrain_model = nat
            yp_nat, _ = predict()
            net.pretrain_model = ref_enc
            mae_ref = All_Metrics(yp_ref, yt, args.mae_thresh, args.mape_thresh)[0]
            mae_nat = All_Metrics(yp_nat, yt, args.mae_thresh, args.mape_thresh)[0]
            mae_ref, mae_nat = float(mae_ref), float(mae_nat)
            self.logger.info("[config5] test MAE reference-encoder {:.6f} native-encoder {:.6f} |dMAE| {:.2e} ; max |d prediction| {:.3e} "
                             "(flow units) over {} test windows".format(mae_ref, mae_nat, abs(mae_ref - mae_nat),
                                                                        (yp_ref - yp_nat).abs().max().item(), yp_ref.shape[0]))

        def ref_enc_args(enc, args):
            import copy
            a = copy.copy(args)
            a.mode = "eval"
            return a

        BT.Trainer.train = train
    sys.argv = ["Run.py"] + argv
    runpy.run_path(os.path.join(model_dir, "Run.py"), run_name="__main__")


if __name__ == "__main__":
    main()
