"""Not a test: prints every parity number in one GPU call (python tests/debug_gpu.py)."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import test_gpu_parity as T
from _util import ENC_CASES, LAYER_CASES


def main():
    print(torch.cuda.get_device_name(0))
    for name in ENC_CASES:
        try:
            T.test_encoder_vs_reference_golden(name)
            print("PASS", name)
        except BaseException as e:  # noqa: BLE001
            print("FAIL", name, str(e)[:3000])
            if not isinstance(e, AssertionError):
                traceback.print_exc()
    for name in LAYER_CASES:
        try:
            T.test_layer_vs_reference_golden(name)
            print("PASS", name)
        except BaseException as e:  # noqa: BLE001
            print("FAIL", name, str(e)[:3000])
            if not isinstance(e, AssertionError):
                traceback.print_exc()


if __name__ == "__main__":
    main()
