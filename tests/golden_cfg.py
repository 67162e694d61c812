"""Dataset geometry of the three reference configs (seg3d/utils/config.py:14-15,
configs/waymo_one_sweep_cylinder.yaml:2-4)."""
CFG = {
    'cart': dict(voxel_size=[0.1, 0.1, 0.1], pc_range=[-72, -72, -2, 72, 72, 4.4]),
    'cyl': dict(voxel_size=[0.05, 0.012, 0.1], pc_range=[0, -3.1415926, -2, 75.2, 3.1415926, 5.2]),
}
