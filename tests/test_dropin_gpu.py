"""The reference's own loops run unchanged against the B200 model (GPU only).

`/root/reference` is not present on the GPU box, so the loop bodies below restate, call for call, what
train_prob_unet_model.py does with the model: :83-102 (train), :125-146 (eval), :168-185 (sample)."""
import sys

import pytest
import torch
from torch.utils.data import DataLoader, Dataset

import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


class _Synthetic(Dataset):
    """Stands in for climex2torch: dict items with 'inputs', 'targets', 'timestamps' (climex_utils.py:122-164)."""

    def __init__(self, n, H):
        self.x, self.t = synth.make_inputs(n, H, H, seed=11)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {'inputs': self.x[i], 'targets': self.t[i], 'timestamps': torch.tensor(float(i), dtype=torch.float64)}


def test_module_alias_and_reference_loops():
    import prob_unet_mds_b200.prob_unet as pu
    sys.modules['prob_unet'] = pu
    from prob_unet import ProbabilisticUNet            # the reference's import line (main.py:7)
    torch.manual_seed(0)
    model = ProbabilisticUNet(input_channels=3, num_classes=3, latent_dim=6, num_filters=[64, 128, 256, 512]).to(DEV)
    model.load_state_dict(synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0))
    loader = DataLoader(_Synthetic(8, 32), batch_size=4, shuffle=False, num_workers=0)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4)       # main.py:95

    # train_probunet_step body (train_prob_unet_model.py:79-102)
    model.train()
    losses = []
    for epoch in range(3):
        for batch in loader:
            inputs = batch['inputs'].to(DEV)
            targets = batch['targets'].to(DEV)
            batch['timestamps'].unsqueeze(dim=1).to(DEV)
            optimizer.zero_grad()
            loss, recon_loss, kl_div = model.elbo(inputs, targets)
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
            assert recon_loss.item() > 0 and kl_div.item() >= 0
    assert all(torch.isfinite(torch.tensor(losses)))
    assert sum(losses[-2:]) < sum(losses[:2])          # it trains

    # eval_probunet_model body (:125-146)
    model.eval()
    with torch.no_grad():
        for batch in loader:
            loss, recon_loss, kl_div = model.elbo(batch['inputs'].to(DEV), batch['targets'].to(DEV))
            assert torch.isfinite(loss)

    # sample_probunet_model body (:168-185)
    with torch.no_grad():
        batch = next(iter(loader))
        inputs = batch['inputs'][:2].to(DEV)
        preds = []
        for _ in range(3):
            output = model(inputs, training=False)
            preds.append(output.cpu())
        stacked = torch.stack(preds, dim=1)
    assert stacked.shape == (2, 3, 3, 32, 32)
    assert not torch.equal(stacked[:, 0], stacked[:, 1])   # different latent draws

    # checkpoint interchange: the state_dict loads back bit for bit
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model2 = ProbabilisticUNet(3, 3, latent_dim=6).to(DEV)
    model2.load_state_dict(sd)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, sd[k])


def test_unmodified_reference_loops_drive_the_b200_model():
    """The reference's own train_prob_unet_model.py (staged byte for byte under oracle/_ref by oracle/make_ref.py; absent
    -> skipped) is imported UNMODIFIED and its train_probunet_step / eval_probunet_model / sample_probunet_model run
    against the B200 model.  Only wandb (not installed; logging only, wandb_active=False) is stubbed."""
    import importlib
    import os
    import types
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip('oracle/_ref is not staged (python oracle/make_ref.py needs /root/reference)')
    sys.modules.setdefault('wandb', types.ModuleType('wandb'))
    if make_ref.REF_DST not in sys.path:
        sys.path.insert(0, make_ref.REF_DST)
    tm = importlib.import_module('train_prob_unet_model')
    assert os.path.dirname(os.path.abspath(tm.__file__)) == make_ref.REF_DST
    from prob_unet_mds_b200 import ProbabilisticUNet

    class _Set(_Synthetic):
        # the extra batch keys and the two dataset methods sample_probunet_model uses (climex_utils.py:156-164, :198-211, :364)
        def __getitem__(self, i):
            d = super().__getitem__(i)
            d.update(lrinterp=self.x[i], hr=self.x[i] + self.t[i], stand_stats=torch.zeros(1))
            return d

        def residual_to_hr(self, residual, lrinterp, stand_stats=None):
            return lrinterp + residual

        def plot_sample_batch(self, *a, **k):
            self.plotted = True
            return None, None

    torch.manual_seed(0)
    model = ProbabilisticUNet(input_channels=3, num_classes=3, latent_dim=6, num_filters=[64, 128, 256, 512]).to(DEV)
    model.load_state_dict(synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=0))
    ds = _Set(8, 32)
    loader = DataLoader(ds, batch_size=4, shuffle=False, num_workers=0)
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4)
    means = [tm.train_probunet_step(model, loader, optimizer, epoch, 3, 1, False, DEV) for epoch in range(3)]
    assert all(m == m for m in means) and means[-1] < means[0]          # finite, and it trains
    val = tm.eval_probunet_model(model, loader, False, DEV)
    assert val == val and val > 0
    hr_preds, _ = tm.sample_probunet_model(model, loader, 0, DEV)
    assert hr_preds.shape == (2, 3, 3, 32, 32) and ds.plotted
    assert not torch.equal(hr_preds[:, 0], hr_preds[:, 1])              # different latent draws
