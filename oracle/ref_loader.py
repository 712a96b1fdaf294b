"""TEST INFRASTRUCTURE — load the UNMODIFIED reference fusion.py in the build container.

/root/reference exists only in the build container, never on the GPU box, so this module
is used by oracle/gen_golden.py (which writes tests/golden/) and by the optional
live-reference tests (skipped when /root/reference is absent).  Nothing is copied out of
the reference: it is imported from where it lies.

fusion.py imports visualisation and perception packages at module scope that are not
installed here (plotly, open3d, groundingdino, segment_anything, ...); a
meta-path finder placed LAST hands out MagicMock modules for exactly those roots, so any
package that is installed is used for real.  The Fusion object is built with __new__,
skipping __init__ (which downloads four networks, reference fusion.py:203-303), and
given the attributes Fusion.eval reads: device, dtype, mu, num_cam, H, W, curr_obs_torch.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('D3F_REFERENCE_ROOT', '/root/reference')

_STUB_ROOTS = ('plotly', 'matplotlib', 'mcubes', 'trimesh', 'open3d', 'groundingdino',
               'segment_anything', 'dgl', 'torchvision', 'cv2', 'PIL', 'tqdm', 'sklearn',
               'pytorch3d', 'supervision', 'huggingface_hub', 'scipy', 'gdown', 'kornia',
               'imageio', 'skimage')


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock()
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'fusion.py'))


_ref_module = None


def load_reference():
    """Import /root/reference/fusion.py unmodified; returns the module."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    spec = importlib.util.spec_from_file_location('d3fields_reference_fusion',
                                                  os.path.join(REFERENCE_ROOT, 'fusion.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ref_module = mod
    return mod


def reference_fusion(scene, device: str = 'cpu'):
    """A reference Fusion whose curr_obs_torch holds `scene` (d3fields_b200.scene.Scene)."""
    import torch
    ref = load_reference()
    F = ref.Fusion.__new__(ref.Fusion)
    F.device = device
    F.dtype = torch.float32
    F.mu = 0.02
    F.num_cam = scene.V
    F.H, F.W = scene.H, scene.W
    obs = {
        'pose': torch.from_numpy(scene.pose).to(device),
        'K': torch.from_numpy(scene.K).to(device),
        'depth': torch.from_numpy(scene.depth).to(device),
    }
    for k, v in scene.maps.items():
        obs[k] = torch.from_numpy(v).to(device=device, dtype=torch.float32)
    F.curr_obs_torch = obs
    return F
