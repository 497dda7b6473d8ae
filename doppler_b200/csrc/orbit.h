// orbit.h -- host-side orbit propagation for `doppler track`: TLE parsing, SGP4 (near-earth) and the
// observer's range rate.  Replaces what the reference gets from crate gpredict 0.2.6 -> C
// libgpredict (/root/reference/src/main.rs:141,149,162-163,170-173; neither is under
// /root/reference nor installed here): Tle::from_file, Predict::new, Predict::update and
// the fields sat.{range_rate_km_sec, az_deg, el_deg, range_km}.
//
// Written from the published algorithm (Hoots & Roehrich, Spacetrack Report No. 3, 1980): SGP4 for
// near-earth element sets and SDP4 (lunar-solar secular and periodic terms, 12 h / 24 h geopotential
// resonance) for periods >= 225 min, selected as libgpredict selects them.  Both are checked against
// that report's own verification cases (tests/test_orbit.py; the resonance terms are exercised by no
// published case and are transcription-checked only).  PARITY UNPINNED against libgpredict itself:
// no copy of it, no TLE file and no reference output exist offline.  Two constant sets are carried:
// kGpredict (default) -- the values libgpredict's sgp4sdp4.h is known to use (WGS-84 radius and
// flattening, rounded qoms2t / s / earth-rotation rate), recalled, not verifiable offline -- and
// kReport3, the strict WGS-72 set of the report.  A sub-hertz difference in the Doppler changes every
// output sample, so `doppler track --tlefile` is functionally equivalent to the reference, NOT
// byte-identical; byte parity is claimed (and tested) for const mode and --doppler-table replay only.
// Pure host code, double precision; it runs once per second of stream, far off the hot path.
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

// Physical constants of the propagator and the observer model.
struct Constants {
    double xkmper;    // equatorial radius, km
    double f;         // flattening
    double qoms2t;    // ((120 - 78) / xkmper)^4
    double s;         // 1 + 78 / xkmper
    double mfactor;   // earth rotation, rad / s (observer velocity)
    const char* name;
};
enum ConstantSet { kGpredict = 0, kReport3 = 1 };
const Constants& constants(int which);
// Process-wide choice for trackers created afterwards (env DOPPLER_B200_ORBIT_CONSTANTS=gpredict|wgs72 at first use).
int default_constant_set();
void set_default_constant_set(int which);

// Deep-space state (SDP4): lunar-solar secular rates and periodic coefficients, resonance integrator.
struct DeepSpace {
    double thgr, xnq, xqncl, omegaq, zmol, zmos, savtsn;
    double ee2, e3, xi2, xi3, xl2, xl3, xl4, xgh2, xgh3, xgh4, xh2, xh3;
    double sse, ssi, ssg, ssh, ssl, se2, si2, sl2, sgh2, sh2, se3, si3, sl3, sgh3, sh3, sl4, sgh4;
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433, del1, del2, del3, fasx2, fasx4, fasx6;
    double xlamo, xfact, omgdt, siniq, cosiq;
    int iresfl, isynfl;
};

class Sgp4 {
public:
    // false (with *err) for unphysical elements.  Element sets with a period >= 225 min take the deep-space model.
    bool init(const Tle& tle, std::string* err, int constant_set = -1);
    // tsince: minutes since the TLE epoch.  pos in km, vel in km/s, TEME frame.
    void propagate(double tsince_min, Vec3* pos_km, Vec3* vel_km_s) const;
    bool deep_space() const { return deep_; }
    const Constants& consts() const { return *k_; }

private:
    void propagate_deep(double tsince_min, Vec3* pos_km, Vec3* vel_km_s) const;
    void deep_init(double eosq, double sinio, double cosio, double betao, double theta2, double sing, double cosg, double betao2,
                   double xmdot, double omgdot, double xnodot, double epoch_jd);
    void deep_secular(double t, double* xll, double* omgasm, double* xnodes, double* em, double* xinc, double* xn) const;
    void deep_periodic(double t, double* em, double* xinc, double* omgasm, double* xnodes, double* xll) const;
    void finish(double a, double e, double omega, double xnode, double xl, double xinc, double xn_unused, Vec3* pos, Vec3* vel) const;

    const Constants* k_ = nullptr;
    bool isimp_ = false, deep_ = false;
    double xmo_, xnodeo_, omegao_, eo_, xincl_, bstar_;
    double aodp_, xnodp_, cosio_, sinio_, x3thm1_, x1mth2_, x7thm1_;
    double c1_, c4_, c5_, d2_, d3_, d4_, eta_, delmo_, sinmo_;
    double xmdot_, omgdot_, xnodot_, omgcof_, xmcof_, xnodcf_, t2cof_, t3cof_, t4cof_, t5cof_, xlcof_, aycof_;
    DeepSpace ds_;
};

struct Observation {
    double az_deg, el_deg, range_km, range_rate_km_s;
};

// Observer on the constant set's ellipsoid; (pos, vel) of the satellite in TEME at Julian date jd.
Observation observe(const Vec3& pos_km, const Vec3& vel_km_s, double jd, double lat_deg, double lon_deg, double alt_m,
                    const Constants& k);

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
    bool deep_space() const { return sgp4_.deep_space(); }
    const Constants& consts() const { return sgp4_.consts(); }

private:
    Tle tle_;
    Sgp4 sgp4_;
    double lat_ = 0, lon_ = 0, alt_ = 0, epoch_jd_ = 0;
    bool have_cache_ = false;
    double cache_t_ = 0;
    Observation cache_{};
};

}  // namespace dorbit
