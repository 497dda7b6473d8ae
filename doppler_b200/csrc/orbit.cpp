// orbit.cpp -- see orbit.h.  SGP4 after Spacetrack Report No. 3 (WGS-72), observer geometry on
// the same ellipsoid.  Variable names follow the report so the equations can be checked against it.
#include "orbit.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>

namespace dorbit {

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kDeg = kPi / 180.0;
constexpr double kAe = 1.0;
constexpr double kTothrd = 2.0 / 3.0;
constexpr double kXkmper = 6378.135;             // WGS-72 equatorial radius, km
constexpr double kF = 1.0 / 298.26;              // WGS-72 flattening
constexpr double kXke = 0.0743669161;            // sqrt(GM) in (earth radii)^1.5 / min
constexpr double kCk2 = 5.413079e-4;             // J2 / 2
constexpr double kCk4 = 6.209887e-7;             // -3 J4 / 8
constexpr double kXj3 = -2.53881e-6;             // J3
constexpr double kQoms2t = 1.880279159015270643865e-9;   // ((120 - 78) / xkmper)^4
constexpr double kS = 1.0122292801892716;        // ae + 78 / xkmper
constexpr double kE6a = 1.0e-6;
constexpr double kMinPerDay = 1440.0;
constexpr double kSecPerDay = 86400.0;
constexpr double kOmegaE = 1.00273790934;        // earth rotations per sidereal day

double fmod2p(double x)
{
    double r = fmod(x, kTwoPi);
    if (r < 0) r += kTwoPi;
    return r;
}

std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
    return s.substr(a, b - a);
}

bool checksum_ok(const std::string& line)
{
    if (line.size() < 69) return false;
    int sum = 0;
    for (int i = 0; i < 68; i++) {
        const char c = line[i];
        if (c >= '0' && c <= '9') sum += c - '0';
        else if (c == '-') sum += 1;
    }
    return line[68] >= '0' && line[68] <= '9' && sum % 10 == line[68] - '0';
}

double field(const std::string& line, int col0, int col1)   // 1-based inclusive columns
{
    return atof(line.substr(col0 - 1, col1 - col0 + 1).c_str());
}

// "-11606-4" -> -0.11606e-4 (implied leading decimal point, exponent without 'e')
double field_exp(const std::string& line, int col0, int col1)
{
    std::string f = line.substr(col0 - 1, col1 - col0 + 1);
    std::string mant, expo;
    size_t i = 0;
    while (i < f.size() && f[i] == ' ') i++;
    double sign = 1.0;
    if (i < f.size() && (f[i] == '-' || f[i] == '+')) {
        if (f[i] == '-') sign = -1.0;
        i++;
    }
    while (i < f.size() && f[i] >= '0' && f[i] <= '9') mant += f[i++];
    while (i < f.size()) expo += f[i++];
    if (mant.empty()) return 0.0;
    const double m = atof(("0." + mant).c_str());
    const int e = expo.empty() ? 0 : atoi(expo.c_str());
    return sign * m * pow(10.0, e);
}

double julian_date_of_year(int year)
{
    year -= 1;
    const long a = year / 100;
    const long b = 2 - a + a / 4;
    return floor(365.25 * year) + floor(30.6001 * 14) + 1720994.5 + (double)b;
}

// Greenwich mean sidereal time (rad) at Julian date jd (IAU 1982 expression, as used with SGP4)
double theta_g(double jd)
{
    const double ut = (jd + 0.5) - floor(jd + 0.5);
    const double jd0 = jd - ut;
    const double tu = (jd0 - 2451545.0) / 36525.0;
    double gmst = 24110.54841 + tu * (8640184.812866 + tu * (0.093104 - tu * 6.2e-6));
    gmst = fmod(gmst + kSecPerDay * kOmegaE * ut, kSecPerDay);
    if (gmst < 0) gmst += kSecPerDay;
    return kTwoPi * gmst / kSecPerDay;
}

}  // namespace

double Tle::epoch_jd() const { return julian_date_of_year(epoch_year) + epoch_day; }

double unix_to_jd(double unix_seconds) { return unix_seconds / kSecPerDay + 2440587.5; }

bool parse_tle_lines(const std::string& name, const std::string& l1, const std::string& l2, Tle* out, std::string* err)
{
    if (l1.size() < 69 || l2.size() < 69 || l1[0] != '1' || l2[0] != '2') {
        if (err) *err = "TLE: malformed element lines for '" + name + "'";
        return false;
    }
    if (!checksum_ok(l1) || !checksum_ok(l2)) {
        if (err) *err = "TLE: checksum mismatch in element set '" + name + "'";
        return false;
    }
    Tle t;
    t.name = name;
    t.catnr = (int)field(l1, 3, 7);
    const int yy = (int)field(l1, 19, 20);
    t.epoch_year = yy < 57 ? 2000 + yy : 1900 + yy;
    t.epoch_day = field(l1, 21, 32);
    t.bstar = field_exp(l1, 54, 61);
    t.incl_deg = field(l2, 9, 16);
    t.raan_deg = field(l2, 18, 25);
    t.ecc = atof(("0." + trim(l2.substr(26, 7))).c_str());
    t.argp_deg = field(l2, 35, 42);
    t.mean_anom_deg = field(l2, 44, 51);
    t.mean_motion_rev_day = field(l2, 53, 63);
    if (!(t.mean_motion_rev_day > 0) || !(t.ecc >= 0 && t.ecc < 1)) {
        if (err) *err = "TLE: unphysical elements in '" + name + "'";
        return false;
    }
    *out = t;
    return true;
}

bool tle_from_file(const std::string& path, const std::string& name, Tle* out, std::string* err)
{
    std::ifstream f(path.c_str());
    if (!f) {
        if (err) *err = "TLE: cannot open file '" + path + "'";
        return false;
    }
    const std::string want = trim(name);
    std::string line, l1, l2;
    while (std::getline(f, line)) {
        if (trim(line) != want) continue;
        if (!std::getline(f, l1) || !std::getline(f, l2)) break;
        while (!l1.empty() && (l1.back() == '\r' || l1.back() == '\n')) l1.pop_back();
        while (!l2.empty() && (l2.back() == '\r' || l2.back() == '\n')) l2.pop_back();
        return parse_tle_lines(want, l1, l2, out, err);
    }
    if (err) *err = "TLE: '" + want + "' not found in '" + path + "'";
    return false;
}

bool Sgp4::init(const Tle& tle, std::string* err)
{
    xmo_ = tle.mean_anom_deg * kDeg;
    xnodeo_ = tle.raan_deg * kDeg;
    omegao_ = tle.argp_deg * kDeg;
    xincl_ = tle.incl_deg * kDeg;
    eo_ = tle.ecc;
    bstar_ = tle.bstar / kAe;
    const double xno = tle.mean_motion_rev_day * kTwoPi / kMinPerDay;   // rad / min

    // recover the original mean motion (xnodp) and semimajor axis (aodp) from the input elements
    const double a1 = pow(kXke / xno, kTothrd);
    cosio_ = cos(xincl_);
    const double theta2 = cosio_ * cosio_;
    x3thm1_ = 3.0 * theta2 - 1.0;
    const double eosq = eo_ * eo_;
    const double betao2 = 1.0 - eosq;
    const double betao = sqrt(betao2);
    const double del1 = 1.5 * kCk2 * x3thm1_ / (a1 * a1 * betao * betao2);
    const double ao = a1 * (1.0 - del1 * (0.5 * kTothrd + del1 * (1.0 + 134.0 / 81.0 * del1)));
    const double delo = 1.5 * kCk2 * x3thm1_ / (ao * ao * betao * betao2);
    xnodp_ = xno / (1.0 + delo);
    aodp_ = ao / (1.0 - delo);

    if (kTwoPi / xnodp_ >= 225.0) {
        if (err) *err = "orbit: '" + tle.name + "' is a deep-space object (period >= 225 min); SDP4 is not implemented";
        return false;
    }

    // for perigee below 220 km the equations are truncated to linear variation in sqrt(a) and
    // quadratic variation in mean anomaly; the c3, delta-omega and delta-m terms are dropped
    isimp_ = (aodp_ * (1.0 - eo_) / kAe) < (220.0 / kXkmper + kAe);

    // for perigee below 156 km the values of s and qoms2t are altered
    double s4 = kS, qoms24 = kQoms2t;
    const double perige = (aodp_ * (1.0 - eo_) - kAe) * kXkmper;
    if (perige < 156.0) {
        s4 = perige <= 98.0 ? 20.0 : perige - 78.0;
        qoms24 = pow((120.0 - s4) * kAe / kXkmper, 4.0);
        s4 = s4 / kXkmper + kAe;
    }
    const double pinvsq = 1.0 / (aodp_ * aodp_ * betao2 * betao2);
    const double tsi = 1.0 / (aodp_ - s4);
    eta_ = aodp_ * eo_ * tsi;
    const double etasq = eta_ * eta_;
    const double eeta = eo_ * eta_;
    const double psisq = fabs(1.0 - etasq);
    const double coef = qoms24 * pow(tsi, 4.0);
    const double coef1 = coef / pow(psisq, 3.5);
    const double c2 = coef1 * xnodp_ *
                      (aodp_ * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                       0.75 * kCk2 * tsi / psisq * x3thm1_ * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    c1_ = bstar_ * c2;
    sinio_ = sin(xincl_);
    const double a3ovk2 = -kXj3 / kCk2 * kAe * kAe * kAe;
    const double c3 = eo_ > 1.0e-12 ? coef * tsi * a3ovk2 * xnodp_ * kAe * sinio_ / eo_ : 0.0;
    x1mth2_ = 1.0 - theta2;
    c4_ = 2.0 * xnodp_ * coef1 * aodp_ * betao2 *
          (eta_ * (2.0 + 0.5 * etasq) + eo_ * (0.5 + 2.0 * etasq) -
           2.0 * kCk2 * tsi / (aodp_ * psisq) *
               (-3.0 * x3thm1_ * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                0.75 * x1mth2_ * (2.0 * etasq - eeta * (1.0 + etasq)) * cos(2.0 * omegao_)));
    c5_ = 2.0 * coef1 * aodp_ * betao2 * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    const double theta4 = theta2 * theta2;
    const double temp1 = 3.0 * kCk2 * pinvsq * xnodp_;
    const double temp2 = temp1 * kCk2 * pinvsq;
    const double temp3 = 1.25 * kCk4 * pinvsq * pinvsq * xnodp_;
    xmdot_ = xnodp_ + 0.5 * temp1 * betao * x3thm1_ + 0.0625 * temp2 * betao * (13.0 - 78.0 * theta2 + 137.0 * theta4);
    const double x1m5th = 1.0 - 5.0 * theta2;
    omgdot_ = -0.5 * temp1 * x1m5th + 0.0625 * temp2 * (7.0 - 114.0 * theta2 + 395.0 * theta4) +
              temp3 * (3.0 - 36.0 * theta2 + 49.0 * theta4);
    const double xhdot1 = -temp1 * cosio_;
    xnodot_ = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * theta2) + 2.0 * temp3 * (3.0 - 7.0 * theta2)) * cosio_;
    omgcof_ = bstar_ * c3 * cos(omegao_);
    xmcof_ = eeta != 0.0 ? -kTothrd * coef * bstar_ * kAe / eeta : 0.0;
    xnodcf_ = 3.5 * betao2 * xhdot1 * c1_;
    t2cof_ = 1.5 * c1_;
    xlcof_ = 0.125 * a3ovk2 * sinio_ * (3.0 + 5.0 * cosio_) / (1.0 + cosio_);
    aycof_ = 0.25 * a3ovk2 * sinio_;
    delmo_ = pow(1.0 + eta_ * cos(xmo_), 3.0);
    sinmo_ = sin(xmo_);
    x7thm1_ = 7.0 * theta2 - 1.0;
    d2_ = d3_ = d4_ = t3cof_ = t4cof_ = t5cof_ = 0.0;
    if (!isimp_) {
        const double c1sq = c1_ * c1_;
        d2_ = 4.0 * aodp_ * tsi * c1sq;
        const double temp = d2_ * tsi * c1_ / 3.0;
        d3_ = (17.0 * aodp_ + s4) * temp;
        d4_ = 0.5 * temp * aodp_ * tsi * (221.0 * aodp_ + 31.0 * s4) * c1_;
        t3cof_ = d2_ + 2.0 * c1sq;
        t4cof_ = 0.25 * (3.0 * d3_ + c1_ * (12.0 * d2_ + 10.0 * c1sq));
        t5cof_ = 0.2 * (3.0 * d4_ + 12.0 * c1_ * d3_ + 6.0 * d2_ * d2_ + 15.0 * c1sq * (2.0 * d2_ + c1sq));
    }
    return true;
}

void Sgp4::propagate(double tsince, Vec3* pos, Vec3* vel) const
{
    // secular gravity and atmospheric drag
    const double xmdf = xmo_ + xmdot_ * tsince;
    const double omgadf = omegao_ + omgdot_ * tsince;
    const double xnoddf = xnodeo_ + xnodot_ * tsince;
    double omega = omgadf;
    double xmp = xmdf;
    const double tsq = tsince * tsince;
    const double xnode = xnoddf + xnodcf_ * tsq;
    double tempa = 1.0 - c1_ * tsince;
    double tempe = bstar_ * c4_ * tsince;
    double templ = t2cof_ * tsq;
    if (!isimp_) {
        const double delomg = omgcof_ * tsince;
        const double delm = xmcof_ * (pow(1.0 + eta_ * cos(xmdf), 3.0) - delmo_);
        const double temp = delomg + delm;
        xmp = xmdf + temp;
        omega = omgadf - temp;
        const double tcube = tsq * tsince;
        const double tfour = tsince * tcube;
        tempa = tempa - d2_ * tsq - d3_ * tcube - d4_ * tfour;
        tempe = tempe + bstar_ * c5_ * (sin(xmp) - sinmo_);
        templ = templ + t3cof_ * tcube + tfour * (t4cof_ + tsince * t5cof_);
    }
    const double a = aodp_ * tempa * tempa;
    const double e = eo_ - tempe;
    const double xl = xmp + omega + xnode + xnodp_ * templ;
    const double beta = sqrt(1.0 - e * e);
    const double xn = kXke / pow(a, 1.5);

    // long period periodics
    const double axn = e * cos(omega);
    double temp = 1.0 / (a * beta * beta);
    const double xll = temp * xlcof_ * axn;
    const double aynl = temp * aycof_;
    const double xlt = xl + xll;
    const double ayn = e * sin(omega) + aynl;

    // solve Kepler's equation
    const double capu = fmod2p(xlt - xnode);
    double temp2 = capu;
    double sinepw = 0, cosepw = 0, temp3 = 0, temp4 = 0, temp5 = 0, temp6 = 0;
    for (int i = 0; i < 10; i++) {
        sinepw = sin(temp2);
        cosepw = cos(temp2);
        temp3 = axn * sinepw;
        temp4 = ayn * cosepw;
        temp5 = axn * cosepw;
        temp6 = ayn * sinepw;
        const double epw = (capu - temp4 + temp3 - temp2) / (1.0 - temp5 - temp6) + temp2;
        if (fabs(epw - temp2) <= kE6a) break;
        temp2 = epw;
    }

    // short period preliminary quantities
    const double ecose = temp5 + temp6;
    const double esine = temp3 - temp4;
    const double elsq = axn * axn + ayn * ayn;
    temp = 1.0 - elsq;
    const double pl = a * temp;
    const double r = a * (1.0 - ecose);
    double temp1 = 1.0 / r;
    const double rdot = kXke * sqrt(a) * esine * temp1;
    const double rfdot = kXke * sqrt(pl) * temp1;
    temp2 = a * temp1;
    const double betal = sqrt(temp);
    temp3 = 1.0 / (1.0 + betal);
    const double cosu = temp2 * (cosepw - axn + ayn * esine * temp3);
    const double sinu = temp2 * (sinepw - ayn - axn * esine * temp3);
    const double u = atan2(sinu, cosu);
    const double sin2u = 2.0 * sinu * cosu;
    const double cos2u = 2.0 * cosu * cosu - 1.0;
    temp = 1.0 / pl;
    temp1 = kCk2 * temp;
    temp2 = temp1 * temp;

    // update for short periodics
    const double rk = r * (1.0 - 1.5 * temp2 * betal * x3thm1_) + 0.5 * temp1 * x1mth2_ * cos2u;
    const double uk = u - 0.25 * temp2 * x7thm1_ * sin2u;
    const double xnodek = xnode + 1.5 * temp2 * cosio_ * sin2u;
    const double xinck = xincl_ + 1.5 * temp2 * cosio_ * sinio_ * cos2u;
    const double rdotk = rdot - xn * temp1 * x1mth2_ * sin2u;
    const double rfdotk = rfdot + xn * temp1 * (x1mth2_ * cos2u + 1.5 * x3thm1_);

    // orientation vectors
    const double sinuk = sin(uk), cosuk = cos(uk);
    const double sinik = sin(xinck), cosik = cos(xinck);
    const double sinnok = sin(xnodek), cosnok = cos(xnodek);
    const double xmx = -sinnok * cosik;
    const double xmy = cosnok * cosik;
    const double ux = xmx * sinuk + cosnok * cosuk;
    const double uy = xmy * sinuk + sinnok * cosuk;
    const double uz = sinik * sinuk;
    const double vx = xmx * cosuk - cosnok * sinuk;
    const double vy = xmy * cosuk - sinnok * sinuk;
    const double vz = sinik * cosuk;

    // position (earth radii -> km) and velocity (earth radii / min -> km / s)
    pos->x = rk * ux * kXkmper;
    pos->y = rk * uy * kXkmper;
    pos->z = rk * uz * kXkmper;
    const double vs = kXkmper / 60.0;
    vel->x = (rdotk * ux + rfdotk * vx) * vs;
    vel->y = (rdotk * uy + rfdotk * vy) * vs;
    vel->z = (rdotk * uz + rfdotk * vz) * vs;
}

Observation observe(const Vec3& pos, const Vec3& vel, double jd, double lat_deg, double lon_deg, double alt_m)
{
    const double lat = lat_deg * kDeg, lon = lon_deg * kDeg, alt_km = alt_m / 1000.0;
    // observer position and velocity in the same inertial frame
    const double theta = fmod2p(theta_g(jd) + lon);   // local mean sidereal time
    const double sinlat = sin(lat), coslat = cos(lat);
    const double c = 1.0 / sqrt(1.0 + kF * (kF - 2.0) * sinlat * sinlat);
    const double sq = (1.0 - kF) * (1.0 - kF) * c;
    const double achcp = (kXkmper * c + alt_km) * coslat;
    const Vec3 opos{achcp * cos(theta), achcp * sin(theta), (kXkmper * sq + alt_km) * sinlat};
    const double mfactor = kTwoPi * kOmegaE / kSecPerDay;   // rad / s
    const Vec3 ovel{-mfactor * opos.y, mfactor * opos.x, 0.0};

    const Vec3 rg{pos.x - opos.x, pos.y - opos.y, pos.z - opos.z};
    const Vec3 rv{vel.x - ovel.x, vel.y - ovel.y, vel.z - ovel.z};
    const double range = sqrt(rg.x * rg.x + rg.y * rg.y + rg.z * rg.z);
    Observation ob;
    ob.range_km = range;
    ob.range_rate_km_s = (rg.x * rv.x + rg.y * rv.y + rg.z * rv.z) / range;

    // topocentric horizon (south, east, zenith)
    const double sinth = sin(theta), costh = cos(theta);
    const double top_s = sinlat * costh * rg.x + sinlat * sinth * rg.y - coslat * rg.z;
    const double top_e = -sinth * rg.x + costh * rg.y;
    const double top_z = coslat * costh * rg.x + coslat * sinth * rg.y + sinlat * rg.z;
    double az = atan2(top_e, -top_s);   // from north, eastward
    if (az < 0) az += kTwoPi;
    ob.az_deg = az / kDeg;
    ob.el_deg = asin(top_z / range) / kDeg;
    return ob;
}

bool Tracker::init(const Tle& tle, double lat_deg, double lon_deg, double alt_m, std::string* err)
{
    tle_ = tle;
    lat_ = lat_deg;
    lon_ = lon_deg;
    alt_ = alt_m;
    epoch_jd_ = tle.epoch_jd();
    have_cache_ = false;
    return sgp4_.init(tle, err);
}

bool Tracker::load(const std::string& tlefile, const std::string& tlename, double lat_deg, double lon_deg, double alt_m, std::string* err)
{
    Tle t;
    if (!tle_from_file(tlefile, tlename, &t, err)) return false;
    return init(t, lat_deg, lon_deg, alt_m, err);
}

Observation Tracker::observe(double unix_seconds) const
{
    const double jd = unix_to_jd(unix_seconds);
    const double tsince = (jd - epoch_jd_) * kMinPerDay;
    Vec3 p, v;
    sgp4_.propagate(tsince, &p, &v);
    return dorbit::observe(p, v, jd, lat_, lon_, alt_);
}

Observation Tracker::observe_cached(double unix_seconds)
{
    if (have_cache_ && cache_t_ == unix_seconds) return cache_;
    cache_ = observe(unix_seconds);
    cache_t_ = unix_seconds;
    have_cache_ = true;
    return cache_;
}

}  // namespace dorbit

// ---- C ABI (include/doppler_b200.h): orbit functions for bindings and tests --------------------
#include "../../include/doppler_b200.h"

struct doppler_b200_tracker {
    dorbit::Tracker t;
    std::string err;
};

static thread_local std::string g_tracker_error;   // per calling thread, like errno

extern "C" {

int doppler_b200_tracker_create(const char* tlefile, const char* tlename, double lat_deg, double lon_deg, double alt_m,
                                doppler_b200_tracker** out)
{
    if (!tlefile || !tlename || !out) return DOPPLER_B200_EINVAL;
    *out = nullptr;
    doppler_b200_tracker* tr = new doppler_b200_tracker;
    if (!tr->t.load(tlefile, tlename, lat_deg, lon_deg, alt_m, &tr->err)) {
        g_tracker_error = tr->err;
        delete tr;
        return DOPPLER_B200_EINVAL;
    }
    *out = tr;
    return DOPPLER_B200_OK;
}

int doppler_b200_tracker_create_from_lines(const char* name, const char* line1, const char* line2, double lat_deg, double lon_deg,
                                           double alt_m, doppler_b200_tracker** out)
{
    if (!line1 || !line2 || !out) return DOPPLER_B200_EINVAL;
    *out = nullptr;
    doppler_b200_tracker* tr = new doppler_b200_tracker;
    dorbit::Tle tle;
    if (!dorbit::parse_tle_lines(name ? name : "", line1, line2, &tle, &tr->err) || !tr->t.init(tle, lat_deg, lon_deg, alt_m, &tr->err)) {
        g_tracker_error = tr->err;
        delete tr;
        return DOPPLER_B200_EINVAL;
    }
    *out = tr;
    return DOPPLER_B200_OK;
}

void doppler_b200_tracker_destroy(doppler_b200_tracker* tr) { delete tr; }

const char* doppler_b200_tracker_last_error(void) { return g_tracker_error.c_str(); }

int doppler_b200_tracker_observe(doppler_b200_tracker* tr, double unix_seconds, double* az_deg, double* el_deg, double* range_km,
                                 double* range_rate_km_sec)
{
    if (!tr) return DOPPLER_B200_EINVAL;
    const dorbit::Observation ob = tr->t.observe(unix_seconds);
    if (az_deg) *az_deg = ob.az_deg;
    if (el_deg) *el_deg = ob.el_deg;
    if (range_km) *range_km = ob.range_km;
    if (range_rate_km_sec) *range_rate_km_sec = ob.range_rate_km_s;
    return DOPPLER_B200_OK;
}

int doppler_b200_tracker_teme(doppler_b200_tracker* tr, double minutes_since_epoch, double* pos_km, double* vel_km_s)
{
    if (!tr || !pos_km || !vel_km_s) return DOPPLER_B200_EINVAL;
    // exposes the bare SGP4 state for verification against published test cases
    dorbit::Sgp4 s;
    std::string err;
    if (!s.init(tr->t.tle(), &err)) return DOPPLER_B200_EINVAL;
    dorbit::Vec3 p, v;
    s.propagate(minutes_since_epoch, &p, &v);
    pos_km[0] = p.x, pos_km[1] = p.y, pos_km[2] = p.z;
    vel_km_s[0] = v.x, vel_km_s[1] = v.y, vel_km_s[2] = v.z;
    return DOPPLER_B200_OK;
}

size_t doppler_b200_tracker_doppler_table(doppler_b200_tracker* tr, double start_unix_seconds, uint32_t frequency, size_t nsec,
                                          double* doppler_hz_out)
{
    if (!tr || !doppler_hz_out) return 0;
    for (size_t s = 0; s < nsec; s++) {
        const dorbit::Observation ob = tr->t.observe(start_unix_seconds + (double)s);
        doppler_hz_out[s] = doppler_b200_doppler_hz(ob.range_rate_km_s, frequency);
    }
    return nsec;
}

}  // extern "C"
