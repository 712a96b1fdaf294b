"""d3fields_b200 — B200-native field query (Fusion.eval / batch_eval) of WangYixuan12/d3fields.

Public surface (mirrors reference fusion.py, see d3fields_b200/fusion.py):
    Fusion, create_init_grid, project_points_coords, interpolate_feats
"""
from .fusion import (Fusion, create_init_grid, create_init_grid_device, project_points_coords,  # noqa: F401
                     interpolate_feats)

__all__ = ['Fusion', 'create_init_grid', 'create_init_grid_device', 'project_points_coords', 'interpolate_feats']
