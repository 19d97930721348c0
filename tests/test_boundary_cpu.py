"""Host-side checks of the drop-in boundary that need no GPU: state_dict schema, constructor surface, and that
the product refuses to run without CUDA instead of silently falling back."""
import inspect

import pytest
import torch

import synth


def test_state_dict_schema_matches_reference():
    from prob_unet_mds_b200 import ProbabilisticUNet
    for L in (6, 16):
        m = ProbabilisticUNet(3, 3, latent_dim=L)
        mine = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert mine == synth.load_schema(f'schema_probunet_L{L}.json')
    assert sum(p.numel() for p in m.parameters()) == 104880643


def test_detunet_schema_matches_reference():
    from prob_unet_mds_b200.baseline.deterministic_unet import UNet
    m = UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False)
    mine = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert mine == synth.load_schema('schema_detunet.json')
    assert sum(p.numel() for p in m.parameters()) == 22792579


def test_constructor_and_method_surface():
    from prob_unet_mds_b200 import AxisAlignedConvGaussian, Fcomb, ProbabilisticUNet
    sig = inspect.signature(ProbabilisticUNet.__init__)
    assert list(sig.parameters)[1:] == ['input_channels', 'num_classes', 'latent_dim', 'num_filters', 'beta']
    assert sig.parameters['latent_dim'].default == 6 and sig.parameters['beta'].default == 1.0
    assert list(inspect.signature(ProbabilisticUNet.forward).parameters)[1:] == ['x', 'target', 'training']
    assert list(inspect.signature(ProbabilisticUNet.elbo).parameters)[1:] == ['x', 'target']
    m = ProbabilisticUNet(3, 3, latent_dim=6)
    for attr in ('unet', 'prior', 'posterior', 'fcomb', 'beta', 'latent_dim'):
        assert hasattr(m, attr)
    assert isinstance(m.prior, AxisAlignedConvGaussian) and isinstance(m.fcomb, Fcomb)
    # zero-initialised tensors of the reference (networks.py:152,162,298)
    sd = m.state_dict()
    for k in ('unet.out_conv.weight', 'unet.enc.64x64_block0.conv1.weight', 'unet.dec.8x8_in0.proj.weight'):
        assert sd[k].abs().max() == 0
    assert sd['unet.enc.64x64_block0.conv0.weight'].abs().max() > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_no_cpu_fallback():
    from prob_unet_mds_b200 import ProbabilisticUNet
    m = ProbabilisticUNet(3, 3, latent_dim=6)
    x, t = synth.make_inputs(1, 32, 32)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m.elbo(x, t)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(x, training=False)


def test_checkpoint_roundtrip_in_reference_format(tmp_path):
    """{name}.pt is a plain state_dict and {name}_optimizer.pt a plain optimizer state_dict, exactly what
    baseline/main.py:108-109 writes; both load back, and a reference-style file (bare torch.save of a state_dict) loads."""
    from prob_unet_mds_b200 import ProbabilisticUNet, checkpoint
    m = ProbabilisticUNet(3, 3, latent_dim=6)
    m.load_state_dict(synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=3))
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    for p in list(m.parameters())[:5]:
        p.grad = torch.ones_like(p)
    opt.step()
    checkpoint.save(str(tmp_path), 'probunet', m, opt, epoch=7, global_step=123)
    raw = torch.load(tmp_path / 'probunet.pt')
    assert [(k, tuple(v.shape)) for k, v in raw.items()] == synth.load_schema('schema_probunet_L6.json')
    m2 = ProbabilisticUNet(3, 3, latent_dim=6)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-4)
    state = checkpoint.load(str(tmp_path), 'probunet', m2, opt2)
    assert state['epoch'] == 7 and state['global_step'] == 123
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()['state'], opt2.state_dict()['state']
    assert s1.keys() == s2.keys() and all(torch.equal(s1[i]['exp_avg'], s2[i]['exp_avg']) for i in s1)
    # a checkpoint written the reference's way
    torch.save(synth.make_weights(synth.load_schema('schema_probunet_L6.json'), seed=4), tmp_path / 'ref.pt')
    assert checkpoint.load(str(tmp_path), 'ref', m2) == {}
