"""Checkpoint I/O in the reference's format.

The reference writes two files per epoch (``baseline/main.py:108-109``):
``torch.save(model.state_dict(), f"{dir}/{name}.pt")`` and ``torch.save(optimizer.state_dict(), f"{dir}/{name}_optimizer.pt")``;
the Prob-UNet driver (``main.py``) never saves anything.  ``save`` / ``load`` keep exactly that layout (so a checkpoint
written by either code base loads in the other: same keys, shapes and dtypes) and add an optional ``{name}_resume.pt``
with what a resumed run needs besides the tensors (epoch / step counters, RNG state)."""
import os

import torch


def _paths(directory, name):
    return (os.path.join(directory, f'{name}.pt'), os.path.join(directory, f'{name}_optimizer.pt'),
            os.path.join(directory, f'{name}_resume.pt'))


def save(directory, name, model, optimizer=None, **resume_state):
    """Writes {name}.pt (+ {name}_optimizer.pt, + {name}_resume.pt when keyword state such as epoch=3 is given)."""
    os.makedirs(directory, exist_ok=True)
    model_path, opt_path, resume_path = _paths(directory, name)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, model_path)
    if optimizer is not None:
        torch.save(optimizer.state_dict(), opt_path)
    if resume_state:
        state = dict(resume_state)
        state['torch_rng_state'] = torch.get_rng_state()
        if torch.cuda.is_available():
            state['cuda_rng_state_all'] = torch.cuda.get_rng_state_all()
        torch.save(state, resume_path)
    return model_path


def load(directory, name, model, optimizer=None, map_location=None, restore_rng=False):
    """Loads what ``save`` (or the reference) wrote; returns the resume dictionary ({} if there is none)."""
    model_path, opt_path, resume_path = _paths(directory, name)
    sd = torch.load(model_path, map_location=map_location or 'cpu')
    if isinstance(sd, dict) and 'state_dict' in sd and not any(torch.is_tensor(v) for v in sd.values()):
        sd = sd['state_dict']
    model.load_state_dict(sd)
    if optimizer is not None and os.path.exists(opt_path):
        optimizer.load_state_dict(torch.load(opt_path, map_location=map_location or 'cpu'))
    state = {}
    if os.path.exists(resume_path):
        state = torch.load(resume_path, map_location='cpu', weights_only=False)
        if restore_rng:
            torch.set_rng_state(state['torch_rng_state'])
            if torch.cuda.is_available() and 'cuda_rng_state_all' in state:
                torch.cuda.set_rng_state_all(state['cuda_rng_state_all'])
    return state
