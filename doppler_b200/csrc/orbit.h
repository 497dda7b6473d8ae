// orbit.h -- host-side orbit propagation for `doppler track`: TLE parsing, SGP4 (near-earth) and the
// observer's range rate.  Replaces what the reference gets from crate gpredict 0.2.6 -> C
// libgpredict (/root/reference/src/main.rs:141,149,162-163,170-173; neither is under
// /root/reference nor installed here): Tle::from_file, Predict::new, Predict::update and
// the fields sat.{range_rate_km_sec, az_deg, el_deg, range_km}.
//
// Written from the published algorithm (Hoots & Roehrich, Spacetrack Report No. 3, 1980: SGP4
// with WGS-72 constants) and checked against that report's own verification case
// (tests/test_orbit.py).  PARITY UNPINNED against libgpredict itself: no copy of it, no TLE file
// and no reference output exist offline; deep-space (SDP4, period >= 225 min) element sets are
// rejected rather than approximated.  Pure host code, double precision; it runs once per second of
// stream, far off the hot path.
#pragma once
#include <stdint.h>

#include <string>

namespace dorbit {

struct Tle {
    std::string name;
    int catnr = 0;
    int epoch_year = 0;        // four digits
    double epoch_day = 0;      // day of year, fractional (1.0 = Jan 1 00:00)
    double bstar = 0;          // 1 / earth radii
    double incl_deg = 0, raan_deg = 0, ecc = 0, argp_deg = 0, mean_anom_deg = 0;
    double mean_motion_rev_day = 0;
    double epoch_jd() const;
};

// Parses the two element lines (69 columns each, checksum verified).  Returns false and sets *err.
bool parse_tle_lines(const std::string& name, const std::string& l1, const std::string& l2, Tle* out, std::string* err);

// Tle::from_file(name, path): the first element set in `path` whose name line equals `name`
// (surrounding blanks ignored).
bool tle_from_file(const std::string& path, const std::string& name, Tle* out, std::string* err);

struct Vec3 {
    double x, y, z;
};

class Sgp4 {
public:
    // false (with *err) for deep-space element sets or unphysical elements
    bool init(const Tle& tle, std::string* err);
    // tsince: minutes since the TLE epoch.  pos in km, vel in km/s, TEME frame.
    void propagate(double tsince_min, Vec3* pos_km, Vec3* vel_km_s) const;

private:
    bool isimp_ = false;
    double xmo_, xnodeo_, omegao_, eo_, xincl_, bstar_;
    double aodp_, xnodp_, cosio_, sinio_, x3thm1_, x1mth2_, x7thm1_;
    double c1_, c4_, c5_, d2_, d3_, d4_, eta_, delmo_, sinmo_;
    double xmdot_, omgdot_, xnodot_, omgcof_, xmcof_, xnodcf_, t2cof_, t3cof_, t4cof_, t5cof_, xlcof_, aycof_;
};

struct Observation {
    double az_deg, el_deg, range_km, range_rate_km_s;
};

// Observer on the WGS-72 ellipsoid; (pos, vel) of the satellite in TEME at Julian date jd.
Observation observe(const Vec3& pos_km, const Vec3& vel_km_s, double jd, double lat_deg, double lon_deg, double alt_m);

double unix_to_jd(double unix_seconds);

// Predict::new + Predict::update in one object.
class Tracker {
public:
    bool load(const std::string& tlefile, const std::string& tlename, double lat_deg, double lon_deg, double alt_m, std::string* err);
    bool init(const Tle& tle, double lat_deg, double lon_deg, double alt_m, std::string* err);
    Observation observe(double unix_seconds) const;
    // Same, memoising the last whole-second query (the replay driver asks once per 8 KiB block
    // but its clock only ticks in whole seconds, main.rs:166).
    Observation observe_cached(double unix_seconds);
    const Tle& tle() const { return tle_; }

private:
    Tle tle_;
    Sgp4 sgp4_;
    double lat_ = 0, lon_ = 0, alt_ = 0, epoch_jd_ = 0;
    bool have_cache_ = false;
    double cache_t_ = 0;
    Observation cache_{};
};

}  // namespace dorbit
