# Packaging shim only (the upstream tree ships no setup.py/pyproject); sources are untouched.
from setuptools import setup
setup(
    name="mine_reference", version="0.0.0",
    py_modules=["synthesis_task", "train", "utils"],
    packages=["network", "network.monodepth2", "operations", "input_pipelines", "input_pipelines.llff",
              "input_pipelines.llff.misc", "visualizations", "configs"],
    package_data={"configs": ["*.yaml"], "visualizations": ["*.jpg"]},
)
