"""``bitblas.testing`` (bitblas/testing/__init__.py:12-91): ``main`` and the tolerance function the reference's
operator tests use."""
import inspect
import sys

import torch


def main():
    import pytest
    test_file = inspect.getsourcefile(sys._getframe(1))
    sys.exit(pytest.main([test_file] + sys.argv[1:]))


def torch_assert_close(tensor_a, tensor_b, rtol=1e-2, atol=1e-3, max_mismatched_ratio=0.001, verbose=False):
    diff = torch.abs(tensor_a - tensor_b)
    max_diff = atol + rtol * torch.abs(tensor_b)
    mismatched = diff > max_diff
    num_mismatched = mismatched.sum().item()
    total_elements = tensor_a.numel()
    max_allowed_mismatched = int(total_elements * max_mismatched_ratio)
    if verbose:
        print(f"Number of mismatched elements: {num_mismatched} / {total_elements} (allowed: {max_allowed_mismatched})")
    if num_mismatched > max_allowed_mismatched:
        raise AssertionError(
            f"Too many mismatched elements: {num_mismatched} > {max_allowed_mismatched} "
            f"({max_mismatched_ratio * 100:.2f}% allowed, but get {num_mismatched / total_elements * 100:.2f}%). "
            f"Greatest absolute difference: {diff.max().item()}, "
            f"Greatest relative difference: {(diff / (torch.abs(tensor_b) + 1e-12)).max().item()}.")
    return True
