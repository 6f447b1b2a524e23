import sys, torch
sys.path.insert(0, ".")
from scripts_gemm_probe import run
