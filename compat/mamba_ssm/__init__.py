"""`mamba_ssm` import paths of the reference models (README.md:6 pins mamba_ssm 2.2.2), served by libb200ssm.

Only the names the reference model files import exist (MedMamba.py:14; SSD/MedSSD.py:32-42; CNN_Mamba.py:24-34;
CrossMamba/CrossMamba_fusion_2b2.py:36-46; MedSSD_kan/*.py, medmamba_kan/*.py the same list)."""
__version__ = "2.2.2+b200"
