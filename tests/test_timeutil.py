"""Start time -> (JD, GMST) helper: pinned against published known answers (SOFA's own test
values for era00 / gmst06, Meeus' Julian-date examples).  Parity with astropy itself -- which the reference app
calls (app.py:107-109) and which is neither vendored nor installed -- is UNPINNED and said so in the module."""
import math

import pytest

from reentry_old_b200 import timeutil as T


def test_julian_date_known_answers():
    assert T.julian_date(2024, 1, 1) == 2460310.5                      # Time('2024-01-01 00:00:00').jd, SURVEY 8(c)
    assert T.julian_date(2000, 1, 1, 12) == 2451545.0                  # J2000.0
    assert T.julian_date(1957, 10, 4, 19, 26, 24) == pytest.approx(2436116.31, abs=1e-9)   # Meeus ex. 7.a
    assert T.julian_date(1987, 1, 27) == 2446822.5                     # Meeus ch. 7
    assert T.julian_date(1600, 12, 31) == 2305812.5
    assert T.julian_date(2024, 2, 29, 6) - T.julian_date(2024, 3, 1, 6) == -1.0


def test_era_and_gmst_sofa_values():
    # t_sofa_c.c: iauEra00(2400000.5, 54388.0) and iauGmst06(2400000.5, 53736.0, 2400000.5, 53736.0)
    assert T.earth_rotation_angle(2400000.5 + 54388.0) == pytest.approx(0.4022837240028158102, abs=1e-12)
    assert T.gmst_iau2006(2400000.5 + 53736.0, 2400000.5 + 53736.0) == pytest.approx(1.754174971870091203, abs=1e-12)
    # ERA at J2000.0 UT1 is the defining constant 2 pi 0.7790572732640
    assert T.earth_rotation_angle(2451545.0) == pytest.approx(2 * math.pi * 0.7790572732640, abs=1e-13)


def test_epoch_and_gmst_scales():
    jd, g = T.epoch_and_gmst(2024, 1, 1, scale="tdb")
    assert jd == 2460310.5 and 0.0 <= g < 2 * math.pi
    # the same instant given in UTC is 69.184 s earlier on the TT axis: GMST differs by the sidereal rate x 69.184 s
    _, g_utc = T.epoch_and_gmst(2024, 1, 1, scale="utc")
    rate = 2 * math.pi * 1.00273781191135448 / 86400.0
    assert (g_utc - g) % (2 * math.pi) == pytest.approx(rate * 69.184, abs=1e-8)   # (+ 69 s of precession in RA: 1.5e-9)
    # one sidereal day later GMST is back
    _, g2 = T.epoch_and_gmst(2024, 1, 1, 23, 56, 4.0905, scale="tdb")
    assert min((g2 - g) % (2 * math.pi), (g - g2) % (2 * math.pi)) < 2e-6
    assert T.tai_minus_utc(T.julian_date(2016, 12, 31)) == 36 and T.tai_minus_utc(T.julian_date(2017, 1, 1)) == 37
    with pytest.raises(ValueError):
        T.epoch_and_gmst(2024, 1, 1, scale="tai")
