"""CPU runs of ``late_checks`` (reference paths of the ops: validates the autograd glue, layouts and index maths)."""
import pytest

import late_checks as L


@pytest.mark.parametrize("k,n,h,w,cin,cout", [(3, 4, 16, 8, 128, 64), (1, 4, 16, 8, 128, 256), (3, 2, 32, 16, 64, 128),
                                             (1, 2, 8, 4, 64, 64)])
def test_conv_stride2_matches_conv2d(k, n, h, w, cin, cout):
    L.check_conv_stride2("cpu", k, n, h, w, cin, cout)
    L.check_conv_stride2("cpu", k, n, h, w, cin, cout, want_stats=True)


def test_fast_head_last_stride2():
    L.check_fast_head_last_stride2("cpu")
