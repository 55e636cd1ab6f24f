"""Start time -> (epoch JD, GMST) without astropy (last item).

The reference app does this with astropy (app.py:107-109):

    epoch = Time(f"{calendar} {h}:{m}:{s}", format="iso", scale="tdb")
    gmst0 = epoch.sidereal_time("mean", "greenwich").to_value(u.rad)

astropy is a third-party dependency that is not vendored under the reference and is not installed
here, so PARITY WITH ASTROPY IS UNPINNED: the formulas below are the published ones astropy delegates
to (ERFA/SOFA: calendar -> JD `cal2jd`, Earth rotation angle `era00`, GMST `gmst06`) and are pinned
against SOFA's own test values (tests/test_timeutil.py), but UT1-UTC (|.| < 0.9 s -> up to 6.6e-5 rad of
GMST), which astropy takes from IERS tables, is taken as a caller-supplied constant (default 0).
Everything downstream (the propagator, the oracle) takes the numeric (epoch_jd, gmst0) pair.
"""
from __future__ import annotations

import math

TWO_PI = 2.0 * math.pi
ARCSEC = math.pi / (180.0 * 3600.0)
TT_MINUS_TAI = 32.184
# TAI-UTC (leap seconds) since 1972, (JD of the UTC day the value starts, seconds)
_LEAP = [(2441317.5, 10), (2441499.5, 11), (2441683.5, 12), (2442048.5, 13), (2442413.5, 14), (2442778.5, 15),
         (2443144.5, 16), (2443509.5, 17), (2443874.5, 18), (2444239.5, 19), (2444786.5, 20), (2445151.5, 21),
         (2445516.5, 22), (2446247.5, 23), (2447161.5, 24), (2447892.5, 25), (2448257.5, 26), (2448804.5, 27),
         (2449169.5, 28), (2449534.5, 29), (2450083.5, 30), (2450630.5, 31), (2451179.5, 32), (2453736.5, 33),
         (2454832.5, 34), (2456109.5, 35), (2457204.5, 36), (2457754.5, 37)]


def julian_date(year: int, month: int, day: int, hour: int = 0, minute: int = 0, second: float = 0.0) -> float:
    """Julian date of a proleptic-Gregorian calendar instant (the integer algorithm of SOFA cal2jd)."""
    my = (month - 14) // 12 if month - 14 >= 0 else -((14 - month) // 12)
    iypmy = year + my
    mjd0 = ((1461 * (iypmy + 4800)) // 4 + (367 * (month - 2 - 12 * my)) // 12
            - (3 * ((iypmy + 4900) // 100)) // 4 + day - 2432076)
    return 2400000.5 + mjd0 + (hour * 3600 + minute * 60 + second) / 86400.0


def tai_minus_utc(jd_utc: float) -> float:
    s = 0.0
    for start, val in _LEAP:
        if jd_utc >= start:
            s = float(val)
    return s


def earth_rotation_angle(jd_ut1: float) -> float:
    """ERA (IAU 2000), radians in [0, 2 pi): SOFA era00."""
    t = jd_ut1 - 2451545.0
    f = math.fmod(jd_ut1, 1.0)   # fraction of the day plus 0.5 from the JD convention, as era00 splits the date
    theta = TWO_PI * (f + 0.7790572732640 + 0.00273781191135448 * t)
    theta = math.fmod(theta, TWO_PI)
    return theta + TWO_PI if theta < 0 else theta


def gmst_iau2006(jd_ut1: float, jd_tt: float) -> float:
    """Greenwich mean sidereal time consistent with IAU 2006 precession, radians in [0, 2 pi): SOFA gmst06."""
    t = (jd_tt - 2451545.0) / 36525.0
    poly = (0.014506 + (4612.156534 + (1.3915817 + (-0.00000044 + (-0.000029956 + (-0.0000000368) * t) * t) * t) * t) * t)
    g = math.fmod(earth_rotation_angle(jd_ut1) + poly * ARCSEC, TWO_PI)
    return g + TWO_PI if g < 0 else g


def epoch_and_gmst(year, month, day, hour=0, minute=0, second=0.0, scale="tdb", ut1_minus_utc=0.0):
    """(epoch_jd, gmst0_rad) for a calendar instant given on `scale` ('tdb' as app.py:107, 'tt' or 'utc').

    epoch_jd is the JD on the given scale (what `Time(...).jd` returns); GMST is evaluated at the
    corresponding UT1 = UTC + ut1_minus_utc.  TDB is taken equal to TT (|TDB - TT| < 1.7 ms -> 1.2e-7 rad)."""
    jd = julian_date(year, month, day, hour, minute, second)
    if scale in ("tdb", "tt"):
        jd_tt = jd
        # TT -> UTC needs the leap-second count of the UTC day: one fixed-point pass is enough away from a leap second
        jd_utc = jd_tt - (TT_MINUS_TAI + tai_minus_utc(jd_tt - 69.184 / 86400.0)) / 86400.0
    elif scale == "utc":
        jd_utc = jd
        jd_tt = jd_utc + (TT_MINUS_TAI + tai_minus_utc(jd_utc)) / 86400.0
    else:
        raise ValueError("scale must be 'tdb', 'tt' or 'utc'")
    jd_ut1 = jd_utc + ut1_minus_utc / 86400.0
    return jd, gmst_iau2006(jd_ut1, jd_tt)
