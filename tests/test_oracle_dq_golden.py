"""The 23 DualQuaternionTest cases of the reference (test/quaternion_test.cpp:57-462), run against
the CPU oracle.  Expected values and the 1e-4 tolerance (quaternion_test.cpp:40) are the reference's."""
import math

import numpy as np
import pytest

MAXERROR = 1e-4
f32 = np.float32
RAD30 = float(f32(math.pi / 6))
RAD45 = float(f32(math.pi / 4))
RAD60 = float(f32(math.pi / 3))
RAD90 = float(f32(math.pi / 2))


@pytest.fixture(scope="module")
def dqs(oracle):
    o = oracle
    return dict(
        o=o,
        dq90=o.dq_from_euler(RAD90, RAD90, RAD90, 0, 0, 0),
        dq60=o.dq_from_euler(RAD60, RAD60, RAD60, 0, 0, 0),
        dq45=o.dq_from_euler(RAD45, RAD45, RAD45, 0, 0, 0),
        dq30Rot=o.dq_from_euler(RAD30, RAD30, RAD30, 0, 0, 0),
        dq0=o.dq_from_euler(0, 0, 0, 0, 0, 0),
        dq30=o.dq_from_euler(0.0, RAD30, 0.0, 0.0, 0.0, 100.0),
    )


def near(a, b, tol=MAXERROR):
    assert abs(float(a) - float(b)) <= tol, (a, b)


def near_vec(a, b, tol=MAXERROR):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert np.all(np.abs(a - b) <= tol), (a, b)


def test_real(dqs):  # :57-66
    near_vec(dqs["dq45"][:4], [0.8446231020115715, 0.19134170284356308, 0.4619399539487806, 0.19134170284356303])


def test_dual(dqs):  # :71-91
    near_vec(dqs["dq30"][:4], [0.9659, 0.0, 0.2588, 0.0])
    near_vec(dqs["dq30"][4:], [0.0, -12.9409, 0.0, 48.2962])


def test_from_rodrigues(dqs):  # :93-120
    o = dqs["o"]
    dq30New = o.dq_from_euler(0.0, RAD30, 0.0, 0, 0, 0)
    t = [0, 0, 0]
    near_vec(o.dq_from_rodrigues([0.0, 0.267949192431123, 0.0], t)[:4], dq30New[:4])
    near_vec(o.dq_from_rodrigues([0.226540919660986, 0.546918160678027, 0.226540919660986], t)[:4], dqs["dq45"][:4])
    near_vec(o.dq_from_rodrigues([0.0, 1.0, 0.0], t)[:4], dqs["dq90"][:4])


def test_sum(dqs):  # :123-141
    s = dqs["o"].dq_add(dqs["dq45"], dqs["dq30"])
    near_vec(s[:4], [1.8105, 0.1913, 0.7208, 0.1913])
    near_vec(s[4:], [0.0, -12.9410, 0.0, 48.2963])


def test_compose_rotations(dqs):  # :144-157
    o = dqs["o"]
    v = [0, 0, 1]
    v1 = o.dq_transform_vertex(dqs["dq90"], v)
    v2 = o.dq_transform_vertex(dqs["dq90"], v1)
    comp = o.dq_mul(dqs["dq90"], dqs["dq90"])
    near_vec(v2, o.dq_transform_vertex(comp, v))


def test_sum_assign(dqs):  # :160-180
    o = dqs["o"]
    s = o.dq_add(o.dq_from_euler(RAD30, RAD45, RAD30, 30, 20, 10), dqs["dq30"])
    near_vec(s[:4], [1.8536, 0.1353, 0.6778, 0.1353])
    near_vec(s[4:], [-6.8953, -0.3683, 7.5233, 57.6655])


def test_diff(dqs):  # :183-201
    d = dqs["o"].dq_sub(dqs["dq45"], dqs["dq30"])
    near_vec(d[:4], [-0.1213, 0.1913, 0.2031, 0.1913])
    near_vec(d[4:], [0.0, 12.9410, 0.0, -48.2963])


def test_diff_assign(dqs):  # :204-224
    o = dqs["o"]
    d = o.dq_sub(o.dq_from_euler(RAD30, RAD45, RAD30, 30, 20, 10), dqs["dq30"])
    near_vec(d[:4], [-0.0783, 0.1353, 0.1601, 0.1353])
    near_vec(d[4:], [-6.8953, 25.5137, 7.5233, -38.9271])


def test_scale(dqs):  # :227-243 : scaling touches the dual part only
    s = dqs["o"].dq_scale(dqs["dq30"], 0.30)
    assert np.array_equal(s[:4], dqs["dq30"][:4])
    near_vec(s[4:], [0.0, -3.8823, 0.0, 14.4889])


def test_scale_assign(dqs):  # :246-265
    o = dqs["o"]
    a = o.dq_from_euler(RAD30, RAD45, RAD30, 30, 20, 10)
    s = o.dq_scale(a, 0.30)
    assert np.array_equal(s[:4], a[:4])
    near_vec(s[4:], [-2.0686, 3.7718, 2.2570, 2.8108])


def test_mul(dqs):  # :268-286
    m = dqs["o"].dq_mul(dqs["dq30"], dqs["dq45"])
    near_vec(m[:4], [0.6963, 0.2343, 0.6648, 0.1353])
    near_vec(m[4:], [-6.7650, -33.2402, 11.7172, 34.8142])


def test_mul_assign(dqs):  # :289-308
    o = dqs["o"]
    m = o.dq_mul(o.dq_from_euler(RAD30, RAD45, RAD30, 30, 20, 10), dqs["dq30"])
    near_vec(m[:4], [0.7490, 0.0957, 0.6344, 0.1657])
    near_vec(m[4:], [-13.3911, 18.4657, -2.8031, 60.5945])


def test_normalize(dqs):  # :311-330 : only the real part is normalised
    o = dqs["o"]
    n = o.dq_normalize(o.dq_add(dqs["dq45"], dqs["dq30"]))
    near_vec(n[:4], [0.9203, 0.0973, 0.3663, 0.0973])
    near_vec(n[4:], [0.0, -12.9410, 0.0, 48.2963])


def test_normalize_asserts_on_zero_real(dqs):  # dual_quaternion.hpp:141
    with pytest.raises(AssertionError):
        dqs["o"].dq_normalize(np.zeros(8, np.float32))


def test_do_not_transform(dqs):  # :333-340
    near_vec(dqs["o"].dq_transform_vertex(dqs["dq0"], [0, 0, 1]), [0, 0, 1])


def test_rotate(dqs):  # :343-350
    near_vec(dqs["o"].dq_transform_vertex(dqs["dq90"], [0, 0, 1]), [1, 0, 0])


def test_translate(dqs):  # :353-362
    o = dqs["o"]
    near_vec(o.dq_transform_vertex(o.dq_from_euler(0, 0, 0, 1, 0, 0), [0, 0, 1]), [1, 0, 1])


def test_translate_and_rotate(dqs):  # :365-374
    o = dqs["o"]
    near_vec(o.dq_transform_vertex(o.dq_from_euler(RAD90, RAD90, RAD90, 1, 0, 0), [0, 0, 1]), [2, 0, 0])


def test_roll(dqs):  # :377-387
    o = dqs["o"]
    near(o.dq_roll(o.dq_from_euler(0, RAD30, 0, 0, 0, 0)), 0)
    near(o.dq_roll(dqs["dq45"]), RAD45)
    near(o.dq_roll(dqs["dq90"]), RAD90)


def test_pitch(dqs):  # :390-398
    o = dqs["o"]
    near(o.dq_pitch(dqs["dq30"]), RAD30)
    near(o.dq_pitch(dqs["dq45"]), RAD45)
    near(o.dq_pitch(dqs["dq90"]), RAD90)


def test_yaw(dqs):  # :401-411
    o = dqs["o"]
    near(o.dq_yaw(o.dq_from_euler(0, RAD30, 0, 0, 0, 0)), 0)
    near(o.dq_yaw(dqs["dq45"]), RAD45)
    near(o.dq_yaw(dqs["dq90"]), RAD90)


def test_euler_angles(dqs):  # :414-435
    o = dqs["o"]
    near_vec(o.dq_euler_angles(o.dq_from_euler(0, RAD30, 0, 0, 0, 0)), [0, RAD30, 0])
    near_vec(o.dq_euler_angles(dqs["dq45"]), [RAD45, RAD45, RAD45])
    near_vec(o.dq_euler_angles(dqs["dq90"]), [RAD90, RAD90, RAD90])


def test_rodrigues(dqs):  # :438-456
    o = dqs["o"]
    near_vec(o.dq_rodrigues(o.dq_from_euler(0, RAD30, 0, 0, 0, 0)), [0, 0.267949192431123, 0])
    near_vec(o.dq_rodrigues(dqs["dq45"]), [0.226540919660986, 0.546918160678027, 0.226540919660986])
    near_vec(o.dq_rodrigues(dqs["dq90"]), [0, 1, 0])


def test_to_string(dqs):  # :458-462
    assert dqs["o"].dq_to_string(dqs["dq30"]) == "real: (0.965926,0,0.258819,0)\ndual: (0,-12.941,0,48.2963)\n"
