"""The parity rule of the boolean queries, shared by every GPU test (and smoke()).

Bar (BASELINE.json north_star): feasible / visible booleans bit-exact outside a 1e-6 m margin band.  The band is two-sided:

* GPU says colliding, oracle says free     -> legal only if the oracle's clearance (min distance over the enabled pairs,
                                               margins subtracted) is <= 1e-6 m;
* GPU says free, oracle says colliding     -> legal only if the oracle's deepest contact is <= 1e-6 m (threshold minus signed
                                               distance; intersecting triangles are measured by their separating-axis
                                               penetration depth, oracle/kb_oracle.c ko_penetration).  A configuration that
                                               violates its joint limits is never inside the band.

`max_bad` bounds how many in-band mismatches a test tolerates; 0 (the default) asserts exact equality with the oracle."""
import numpy as np

BAND = 1e-6


def band_violations(got, want, Q, orc, include_self=True):
    """(indices of mismatches, messages of those that are outside the band)"""
    got, want = np.asarray(got), np.asarray(want)
    bad = np.nonzero(got != want)[0]
    out = []
    for i in bad:
        if want[i]:      # oracle free, GPU colliding
            d, _ = orc.distance(Q[i], upper_bound=1.0, include_self=include_self)
            if not (d <= BAND):
                out.append("config %d: gpu=%d oracle=%d but the oracle's clearance is %g m (> band)" % (i, got[i], want[i], d))
        else:            # oracle colliding (or limits violated), GPU free
            if not orc.check_joint_limits(Q[i]):
                out.append("config %d: gpu=feasible but the joint limits are violated" % i)
                continue
            p = orc.penetration(Q[i], include_self=include_self)
            if not (p <= BAND):
                out.append("config %d: gpu=%d oracle=%d but the oracle's deepest contact is %g m (> band)" % (i, got[i], want[i], p))
    return bad, out


def assert_bool_parity(got, want, Q, orc, include_self=True, max_bad=0):
    bad, out = band_violations(got, want, Q, orc, include_self)
    assert not out, "; ".join(out[:5]) + (" (+%d more)" % (len(out) - 5) if len(out) > 5 else "")
    assert len(bad) <= max_bad, "%d boolean mismatches inside the band in %d configurations (allowed %d)" % (len(bad), len(got), max_bad)
    return len(bad)


def assert_geom_bool_parity(got, want, ga, Ta, gb, Tb, orc, tol=0.0, max_bad=0):
    """the same rule for an explicit geometry pair at N transform pairs, threshold = margins + tol"""
    got, want = np.asarray(got).astype(bool), np.asarray(want).astype(bool)
    bad = np.nonzero(got != want)[0]
    for i in bad:
        if not want[i]:  # oracle: not within tol; GPU: within -> clearance above the threshold must be <= band
            d = orc.geom_distance(ga, Ta[i], gb, Tb[i])
            assert d - tol <= BAND, "pair %d: gpu hit, oracle distance %g (threshold %g)" % (i, d, tol)
        else:
            p = orc.geom_penetration(ga, Ta[i], gb, Tb[i], tol)
            assert p <= BAND, "pair %d: gpu miss, oracle contact depth %g" % (i, p)
    assert len(bad) <= max_bad, "%d mismatches inside the band (allowed %d)" % (len(bad), max_bad)
    return len(bad)
