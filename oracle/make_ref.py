"""Stages the UNMODIFIED reference modules of the hot path under oracle/_ref/ (git-ignored, travels with gpurun).

    python oracle/make_ref.py          (run in the authoring container, where /root/reference exists;
                                        __graft_entry__.build() calls it)

The reference is a flat directory of Python scripts -- there is nothing to compile and no package to install -- so
"building the reference" is staging the files of the path byte for byte, with a manifest of their SHA-256 digests:

    prob_unet.py, networks.py              the path itself
    train_prob_unet_model.py               its caller: the train / eval / sample loops
    baseline/deterministic_unet.py         config 5

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (`--impl reference`, `cpu_baseline`) and its informational
`gpu_eager_reference` leg may import from oracle/_ref; the product never does.  Nothing under oracle/_ref is committed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
REF_DST = os.path.join(HERE, '_ref')
FILES = ['prob_unet.py', 'networks.py', 'train_prob_unet_model.py', 'baseline/deterministic_unet.py']


def stage(src=REF_SRC, dst=REF_DST):
    """Returns the manifest {relative path: sha256}, or None when the reference is not present (GPU box)."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        s = os.path.join(src, rel)
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'sha256': manifest}, f, indent=1)
    return manifest


def available(dst=REF_DST):
    return all(os.path.exists(os.path.join(dst, rel)) for rel in FILES)


def import_reference(dst=REF_DST):
    """Imports the staged reference modules (`prob_unet`, `networks`) and returns the `prob_unet` module.  wandb (absent
    here, used only for logging in train_prob_unet_model.py) is stubbed by the callers that need that module."""
    if not available(dst):
        raise ImportError(f'{dst} is not staged (run python oracle/make_ref.py where /root/reference exists)')
    for name in ('prob_unet', 'networks'):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, '__file__', '').startswith(dst):
            del sys.modules[name]            # e.g. the product module aliased as `prob_unet` by a drop-in test
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import prob_unet
    return prob_unet


if __name__ == '__main__':
    m = stage()
    print('oracle/_ref:', 'reference not present, nothing staged' if m is None else json.dumps(m, indent=1))
