"""mine_b200: single-image novel view synthesis with continuous-depth multiplane images (the capabilities of
vincentfung13/MINE, reference ``synthesis_task.py`` / ``train.py`` / ``visualizations/image_to_video.py``), built for
NVIDIA B200 (sm_100a): ``task`` (public facade), ``engine`` / ``ops`` (tcgen05 conv engine, fused kernels), ``models``,
``spec`` (PyTorch specification of every op), ``parallel`` (NVLink collectives), ``data``, ``utils``."""
