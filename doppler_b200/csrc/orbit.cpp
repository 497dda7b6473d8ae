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
constexpr double kXke = 0.0743669161;            // sqrt(GM) in (earth radii)^1.5 / min
constexpr double kCk2 = 5.413079e-4;             // J2 / 2
constexpr double kCk4 = 6.209887e-7;             // -3 J4 / 8
constexpr double kXj3 = -2.53881e-6;             // J3
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

// The two constant sets (orbit.h).  kGpredict: what libgpredict's sgp4sdp4.h carries (WGS-84 radius and flattening next to
// the WGS-72 gravity field, qoms2t / s / the rotation rate rounded as printed there) -- recalled, not verifiable offline.
// kReport3: Spacetrack Report No. 3's own WGS-72 values, qoms2t and s at full precision.
const Constants kSets[2] = {
    {6378.137, 3.35281066474748e-3, 1.880279e-9, 1.012229, 7.292115e-5, "gpredict"},
    {6378.135, 1.0 / 298.26, 1.880279159015270643865e-9, 1.0122292801892716, kTwoPi * kOmegaE / kSecPerDay, "wgs72"},
};
int g_default_set = -1;

}  // namespace

const Constants& constants(int which) { return kSets[which == kReport3 ? kReport3 : kGpredict]; }

int default_constant_set()
{
    if (g_default_set < 0) {
        const char* e = getenv("DOPPLER_B200_ORBIT_CONSTANTS");
        g_default_set = (e && (!strcmp(e, "wgs72") || !strcmp(e, "report3"))) ? kReport3 : kGpredict;
    }
    return g_default_set;
}

void set_default_constant_set(int which) { g_default_set = which == kReport3 ? kReport3 : kGpredict; }

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

bool Sgp4::init(const Tle& tle, std::string* err, int constant_set)
{
    k_ = &constants(constant_set < 0 ? default_constant_set() : constant_set);
    const double kXkmper = k_->xkmper, kQoms2t = k_->qoms2t, kS = k_->s;
    (void)err;
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

    // periods of 225 min and more take the deep-space model (libgpredict: select_ephemeris)
    deep_ = kTwoPi / xnodp_ >= 225.0;

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
    if (deep_) {
        deep_init(eosq, sinio_, cosio_, betao, theta2, sin(omegao_), cos(omegao_), betao2, xmdot_, omgdot_, xnodot_, tle.epoch_jd());
        return true;
    }
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
    if (deep_) {
        propagate_deep(tsince, pos, vel);
        return;
    }
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
    finish(a, e, omega, xnode, xl, xincl_, 0.0, pos, vel);
}

// SDP4 (Spacetrack Report No. 3, section 7): the SGP4 secular / drag terms truncated as for low perigees, plus the
// deep-space secular (deep_secular), resonance and lunar-solar periodic (deep_periodic) corrections.
void Sgp4::propagate_deep(double tsince, Vec3* pos, Vec3* vel) const
{
    double xmdf = xmo_ + xmdot_ * tsince;
    double omgadf = omegao_ + omgdot_ * tsince;
    const double xnoddf = xnodeo_ + xnodot_ * tsince;
    const double tsq = tsince * tsince;
    double xnode = xnoddf + xnodcf_ * tsq;
    const double tempa = 1.0 - c1_ * tsince;
    const double tempe = bstar_ * c4_ * tsince;
    const double templ = t2cof_ * tsq;
    double xn = xnodp_, em = eo_, xinc = xincl_;
    deep_secular(tsince, &xmdf, &omgadf, &xnode, &em, &xinc, &xn);
    const double a = pow(kXke / xn, kTothrd) * tempa * tempa;
    double e = em - tempe;
    double xmam = xmdf + xnodp_ * templ;
    deep_periodic(tsince, &e, &xinc, &omgadf, &xnode, &xmam);
    const double xl = xmam + omgadf + xnode;
    finish(a, e, omgadf, xnode, xl, xinc, 0.0, pos, vel);
}

// Long-period periodics, Kepler's equation, short-period periodics and the orientation vectors: common to SGP4 and SDP4
// (the short-period coefficients x3thm1 .. keep their epoch values in both, as in the report).
void Sgp4::finish(double a, double e, double omega, double xnode, double xl, double xinc, double, Vec3* pos, Vec3* vel) const
{
    const double kXkmper = k_->xkmper;
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
    const double xinck = xinc + 1.5 * temp2 * cosio_ * sinio_ * cos2u;
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

// ---- deep space (Spacetrack Report No. 3, subroutine DEEP) ------------------------------------------------------------
namespace {
// lunar-solar and resonance constants of the report
constexpr double kZns = 1.19459e-5, kC1ss = 2.9864797e-6, kZes = 0.01675, kZnl = 1.5835218e-4, kC1l = 4.7968065e-7, kZel = 0.05490;
constexpr double kZcosis = 0.91744867, kZsinis = 0.39785416, kZsings = -0.98088458, kZcosgs = 0.1945905;
constexpr double kQ22 = 1.7891679e-6, kQ31 = 2.1460748e-6, kQ33 = 2.2123015e-7;
constexpr double kG22 = 5.7686396, kG32 = 0.95240898, kG44 = 1.8014998, kG52 = 1.0508330, kG54 = 4.4108898;
constexpr double kRoot22 = 1.7891679e-6, kRoot32 = 3.7393792e-7, kRoot44 = 7.3636953e-9, kRoot52 = 1.1428639e-7, kRoot54 = 2.1765803e-9;
constexpr double kThdt = 4.3752691e-3;   // earth rotation, rad / min
constexpr double kStepp = 720.0, kStepn = -720.0, kStep2 = 259200.0;

double actan(double sinx, double cosx)   // the report's ACTAN: angle in [0, 2 pi)
{
    double a = atan2(sinx, cosx);
    if (a < 0) a += kTwoPi;
    return a;
}
}  // namespace

// DEEP entry DPINIT.
void Sgp4::deep_init(double eosq, double sinio, double cosio, double betao, double theta2, double sing, double cosg, double betao2,
                     double xmdot, double omgdot, double xnodot, double epoch_jd)
{
    DeepSpace& d = ds_;
    memset(&d, 0, sizeof d);
    // Greenwich sidereal angle at epoch, the report's THETAG: days since 1950 Jan 0.0 UT
    const double ds50 = epoch_jd - 2433281.5;
    d.thgr = fmod2p(6.3003880987 * ds50 + 1.72944494);
    const double eq = eo_;
    d.xnq = xnodp_;
    const double aqnv = 1.0 / aodp_;
    d.xqncl = xincl_;
    const double xmao = xmo_;
    const double xpidot = omgdot + xnodot;
    const double sinq = sin(xnodeo_), cosq = cos(xnodeo_);
    d.omegaq = omegao_;
    d.omgdt = omgdot;
    d.siniq = sinio;
    d.cosiq = cosio;
    const double siniq = sinio, cosiq = cosio, eqsq = eosq, rteqsq = betao, cosq2 = theta2, sinomo = sing, cosomo = cosg, bsq = betao2;

    // lunar-solar terms: geometry of the moon's orbit at epoch
    const double day = ds50 + 18261.5;
    const double xnodce = 4.5236020 - 9.2422029e-4 * day;
    const double stem = sin(xnodce), ctem = cos(xnodce);
    const double zcosil = 0.91375164 - 0.03568096 * ctem;
    const double zsinil = sqrt(1.0 - zcosil * zcosil);
    const double zsinhl = 0.089683511 * stem / zsinil;
    const double zcoshl = sqrt(1.0 - zsinhl * zsinhl);
    const double c = 4.7199672 + 0.22997150 * day;
    const double gam = 5.8351514 + 0.0019443680 * day;
    d.zmol = fmod2p(c - gam);
    double zx = 0.39785416 * stem / zsinil;
    const double zy = zcoshl * ctem + 0.91744867 * zsinhl * stem;
    zx = actan(zx, zy);
    zx = gam + zx - xnodce;
    const double zcosgl = cos(zx), zsingl = sin(zx);
    d.zmos = fmod2p(6.2565837 + 0.017201977 * day);

    // solar terms first, then the same with the moon's constants
    d.savtsn = 1.0e20;
    double zcosg = kZcosgs, zsing = kZsings, zcosi = kZcosis, zsini = kZsinis, zcosh = cosq, zsinh = sinq;
    double cc = kC1ss, zn = kZns, ze = kZes;
    const double xnoi = 1.0 / d.xnq;
    double se = 0, si = 0, sl = 0, sgh = 0, sh = 0;
    for (int pass = 0; pass < 2; pass++) {
        const double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
        const double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
        const double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
        const double a8 = zsing * zsini;
        const double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
        const double a10 = zcosg * zsini;
        const double a2 = cosiq * a7 + siniq * a8;
        const double a4 = cosiq * a9 + siniq * a10;
        const double a5 = -siniq * a7 + cosiq * a8;
        const double a6 = -siniq * a9 + cosiq * a10;
        const double x1 = a1 * cosomo + a2 * sinomo;
        const double x2 = a3 * cosomo + a4 * sinomo;
        const double x3 = -a1 * sinomo + a2 * cosomo;
        const double x4 = -a3 * sinomo + a4 * cosomo;
        const double x5 = a5 * sinomo;
        const double x6 = a6 * sinomo;
        const double x7 = a5 * cosomo;
        const double x8 = a6 * cosomo;
        const double z31 = 12.0 * x1 * x1 - 3.0 * x3 * x3;
        const double z32 = 24.0 * x1 * x2 - 6.0 * x3 * x4;
        const double z33 = 12.0 * x2 * x2 - 3.0 * x4 * x4;
        double z1 = 3.0 * (a1 * a1 + a2 * a2) + z31 * eqsq;
        double z2 = 6.0 * (a1 * a3 + a2 * a4) + z32 * eqsq;
        double z3 = 3.0 * (a3 * a3 + a4 * a4) + z33 * eqsq;
        const double z11 = -6.0 * a1 * a5 + eqsq * (-24.0 * x1 * x7 - 6.0 * x3 * x5);
        const double z12 = -6.0 * (a1 * a6 + a3 * a5) + eqsq * (-24.0 * (x2 * x7 + x1 * x8) - 6.0 * (x3 * x6 + x4 * x5));
        const double z13 = -6.0 * a3 * a6 + eqsq * (-24.0 * x2 * x8 - 6.0 * x4 * x6);
        const double z21 = 6.0 * a2 * a5 + eqsq * (24.0 * x1 * x5 - 6.0 * x3 * x7);
        const double z22 = 6.0 * (a4 * a5 + a2 * a6) + eqsq * (24.0 * (x2 * x5 + x1 * x6) - 6.0 * (x4 * x7 + x3 * x8));
        const double z23 = 6.0 * a4 * a6 + eqsq * (24.0 * x2 * x6 - 6.0 * x4 * x8);
        z1 = z1 + z1 + bsq * z31;
        z2 = z2 + z2 + bsq * z32;
        z3 = z3 + z3 + bsq * z33;
        const double s3 = cc * xnoi;
        const double s2 = -0.5 * s3 / rteqsq;
        const double s4 = s3 * rteqsq;
        const double s1 = -15.0 * eq * s4;
        const double s5 = x1 * x3 + x2 * x4;
        const double s6 = x2 * x3 + x1 * x4;
        const double s7 = x2 * x4 - x1 * x3;
        se = s1 * zn * s5;
        si = s2 * zn * (z11 + z13);
        sl = -zn * s3 * (z1 + z3 - 14.0 - 6.0 * eqsq);
        sgh = s4 * zn * (z31 + z33 - 6.0);
        sh = -zn * s2 * (z21 + z23);
        if (d.xqncl < 5.2359877e-2) sh = 0.0;
        d.ee2 = 2.0 * s1 * s6;
        d.e3 = 2.0 * s1 * s7;
        d.xi2 = 2.0 * s2 * z12;
        d.xi3 = 2.0 * s2 * (z13 - z11);
        d.xl2 = -2.0 * s3 * z2;
        d.xl3 = -2.0 * s3 * (z3 - z1);
        d.xl4 = -2.0 * s3 * (-21.0 - 9.0 * eqsq) * ze;
        d.xgh2 = 2.0 * s4 * z32;
        d.xgh3 = 2.0 * s4 * (z33 - z31);
        d.xgh4 = -18.0 * s4 * ze;
        d.xh2 = -2.0 * s2 * z22;
        d.xh3 = -2.0 * s2 * (z23 - z21);
        if (pass == 1) break;
        // keep the solar terms, set up the lunar pass
        d.sse = se;
        d.ssi = si;
        d.ssl = sl;
        d.ssh = sh / siniq;
        d.ssg = sgh - cosiq * d.ssh;
        d.se2 = d.ee2, d.si2 = d.xi2, d.sl2 = d.xl2, d.sgh2 = d.xgh2, d.sh2 = d.xh2;
        d.se3 = d.e3, d.si3 = d.xi3, d.sl3 = d.xl3, d.sgh3 = d.xgh3, d.sh3 = d.xh3;
        d.sl4 = d.xl4, d.sgh4 = d.xgh4;
        zcosg = zcosgl, zsing = zsingl, zcosi = zcosil, zsini = zsinil;
        zcosh = zcoshl * cosq + zsinhl * sinq;
        zsinh = sinq * zcoshl - cosq * zsinhl;
        zn = kZnl, cc = kC1l, ze = kZel;
    }
    d.sse += se;
    d.ssi += si;
    d.ssl += sl;
    d.ssg += sgh - cosiq / siniq * sh;
    d.ssh += sh / siniq;

    // geopotential resonance: 24 h (synchronous) and 12 h orbits
    d.iresfl = d.isynfl = 0;
    double bfact;
    if (d.xnq < 0.0052359877 && d.xnq > 0.0034906585) {
        d.iresfl = d.isynfl = 1;
        const double g200 = 1.0 + eqsq * (-2.5 + 0.8125 * eqsq);
        const double g310 = 1.0 + 2.0 * eqsq;
        const double g300 = 1.0 + eqsq * (-6.0 + 6.60937 * eqsq);
        const double f220 = 0.75 * (1.0 + cosiq) * (1.0 + cosiq);
        const double f311 = 0.9375 * siniq * siniq * (1.0 + 3.0 * cosiq) - 0.75 * (1.0 + cosiq);
        double f330 = 1.0 + cosiq;
        f330 = 1.875 * f330 * f330 * f330;
        d.del1 = 3.0 * d.xnq * d.xnq * aqnv * aqnv;
        d.del2 = 2.0 * d.del1 * f220 * g200 * kQ22;
        d.del3 = 3.0 * d.del1 * f330 * g300 * kQ33 * aqnv;
        d.del1 = d.del1 * f311 * g310 * kQ31 * aqnv;
        d.fasx2 = 0.13130908, d.fasx4 = 2.8843198, d.fasx6 = 0.37448087;
        d.xlamo = xmao + xnodeo_ + omegao_ - d.thgr;
        bfact = xmdot + xpidot - kThdt;
        bfact = bfact + d.ssl + d.ssg + d.ssh;
    } else if (d.xnq < 8.26e-3 || d.xnq > 9.24e-3 || eq < 0.5) {
        return;
    } else {
        d.iresfl = 1;
        const double eoc = eq * eqsq;
        const double g201 = -0.306 - (eq - 0.64) * 0.440;
        double g211, g310, g322, g410, g422, g520, g533, g521, g532;
        if (eq <= 0.65) {
            g211 = 3.616 - 13.247 * eq + 16.290 * eqsq;
            g310 = -19.302 + 117.390 * eq - 228.419 * eqsq + 156.591 * eoc;
            g322 = -18.9068 + 109.7927 * eq - 214.6334 * eqsq + 146.5816 * eoc;
            g410 = -41.122 + 242.694 * eq - 471.094 * eqsq + 313.953 * eoc;
            g422 = -146.407 + 841.880 * eq - 1629.014 * eqsq + 1083.435 * eoc;
            g520 = -532.114 + 3017.977 * eq - 5740.0 * eqsq + 3708.276 * eoc;
        } else {
            g211 = -72.099 + 331.819 * eq - 508.738 * eqsq + 266.724 * eoc;
            g310 = -346.844 + 1582.851 * eq - 2415.925 * eqsq + 1246.113 * eoc;
            g322 = -342.585 + 1554.908 * eq - 2366.899 * eqsq + 1215.972 * eoc;
            g410 = -1052.797 + 4758.686 * eq - 7193.992 * eqsq + 3651.957 * eoc;
            g422 = -3581.69 + 16178.11 * eq - 24462.77 * eqsq + 12422.52 * eoc;
            g520 = eq <= 0.715 ? 1464.74 - 4664.75 * eq + 3763.64 * eqsq : -5149.66 + 29936.92 * eq - 54087.36 * eqsq + 31324.56 * eoc;
        }
        if (eq < 0.7) {
            g533 = -919.2277 + 4988.61 * eq - 9064.77 * eqsq + 5542.21 * eoc;
            g521 = -822.71072 + 4568.6173 * eq - 8491.4146 * eqsq + 5337.524 * eoc;
            g532 = -853.666 + 4690.25 * eq - 8624.77 * eqsq + 5341.4 * eoc;
        } else {
            g533 = -37995.78 + 161616.52 * eq - 229838.2 * eqsq + 109377.94 * eoc;
            g521 = -51752.104 + 218913.95 * eq - 309468.16 * eqsq + 146349.42 * eoc;
            g532 = -40023.88 + 170470.89 * eq - 242699.48 * eqsq + 115605.82 * eoc;
        }
        const double sini2 = siniq * siniq;
        const double f220 = 0.75 * (1.0 + 2.0 * cosiq + cosq2);
        const double f221 = 1.5 * sini2;
        const double f321 = 1.875 * siniq * (1.0 - 2.0 * cosiq - 3.0 * cosq2);
        const double f322 = -1.875 * siniq * (1.0 + 2.0 * cosiq - 3.0 * cosq2);
        const double f441 = 35.0 * sini2 * f220;
        const double f442 = 39.3750 * sini2 * sini2;
        const double f522 = 9.84375 * siniq * (sini2 * (1.0 - 2.0 * cosiq - 5.0 * cosq2) + 0.33333333 * (-2.0 + 4.0 * cosiq + 6.0 * cosq2));
        const double f523 = siniq * (4.92187512 * sini2 * (-2.0 - 4.0 * cosiq + 10.0 * cosq2) + 6.56250012 * (1.0 + 2.0 * cosiq - 3.0 * cosq2));
        const double f542 = 29.53125 * siniq * (2.0 - 8.0 * cosiq + cosq2 * (-12.0 + 8.0 * cosiq + 10.0 * cosq2));
        const double f543 = 29.53125 * siniq * (-2.0 - 8.0 * cosiq + cosq2 * (12.0 + 8.0 * cosiq - 10.0 * cosq2));
        const double xno2 = d.xnq * d.xnq, ainv2 = aqnv * aqnv;
        double temp1 = 3.0 * xno2 * ainv2;
        double temp = temp1 * kRoot22;
        d.d2201 = temp * f220 * g201;
        d.d2211 = temp * f221 * g211;
        temp1 = temp1 * aqnv;
        temp = temp1 * kRoot32;
        d.d3210 = temp * f321 * g310;
        d.d3222 = temp * f322 * g322;
        temp1 = temp1 * aqnv;
        temp = 2.0 * temp1 * kRoot44;
        d.d4410 = temp * f441 * g410;
        d.d4422 = temp * f442 * g422;
        temp1 = temp1 * aqnv;
        temp = temp1 * kRoot52;
        d.d5220 = temp * f522 * g520;
        d.d5232 = temp * f523 * g532;
        temp = 2.0 * temp1 * kRoot54;
        d.d5421 = temp * f542 * g521;
        d.d5433 = temp * f543 * g533;
        d.xlamo = xmao + xnodeo_ + xnodeo_ - d.thgr - d.thgr;
        bfact = xmdot + xnodot + xnodot - kThdt - kThdt;
        bfact = bfact + d.ssl + d.ssh + d.ssh;
    }
    d.xfact = bfact - d.xnq;
}

// DEEP entry DPSEC: lunar-solar secular rates and, for resonant orbits, the numerically integrated mean motion / longitude.
// The report's integrator keeps its state between calls (ATIME, XLI, XNI); this one restarts from the epoch on every call
// -- same steps, same values for any t, no dependence on the order the caller asks in.
void Sgp4::deep_secular(double t, double* xll, double* omgasm, double* xnodes, double* em, double* xinc, double* xn) const
{
    const DeepSpace& d = ds_;
    *xll += d.ssl * t;
    *omgasm += d.ssg * t;
    *xnodes += d.ssh * t;
    *em = eo_ + d.sse * t;
    *xinc = d.xqncl + d.ssi * t;
    if (*xinc < 0.0) {
        *xinc = -*xinc;
        *xnodes += kPi;
        *omgasm -= kPi;
    }
    if (!d.iresfl) return;

    double atime = 0.0, xni = d.xnq, xli = d.xlamo;
    const double delt = t >= 0.0 ? kStepp : kStepn;
    double xndot = 0, xnddt = 0, xldot = 0;
    auto dots = [&]() {
        if (d.isynfl) {
            xndot = d.del1 * sin(xli - d.fasx2) + d.del2 * sin(2.0 * (xli - d.fasx4)) + d.del3 * sin(3.0 * (xli - d.fasx6));
            xnddt = d.del1 * cos(xli - d.fasx2) + 2.0 * d.del2 * cos(2.0 * (xli - d.fasx4)) + 3.0 * d.del3 * cos(3.0 * (xli - d.fasx6));
        } else {
            const double xomi = d.omegaq + d.omgdt * atime;
            const double x2omi = xomi + xomi, x2li = xli + xli;
            xndot = d.d2201 * sin(x2omi + xli - kG22) + d.d2211 * sin(xli - kG22) + d.d3210 * sin(xomi + xli - kG32) +
                    d.d3222 * sin(-xomi + xli - kG32) + d.d4410 * sin(x2omi + x2li - kG44) + d.d4422 * sin(x2li - kG44) +
                    d.d5220 * sin(xomi + xli - kG52) + d.d5232 * sin(-xomi + xli - kG52) + d.d5421 * sin(xomi + x2li - kG54) +
                    d.d5433 * sin(-xomi + x2li - kG54);
            xnddt = d.d2201 * cos(x2omi + xli - kG22) + d.d2211 * cos(xli - kG22) + d.d3210 * cos(xomi + xli - kG32) +
                    d.d3222 * cos(-xomi + xli - kG32) + d.d5220 * cos(xomi + xli - kG52) + d.d5232 * cos(-xomi + xli - kG52) +
                    2.0 * (d.d4410 * cos(x2omi + x2li - kG44) + d.d4422 * cos(x2li - kG44) + d.d5421 * cos(xomi + x2li - kG54) +
                           d.d5433 * cos(-xomi + x2li - kG54));
        }
        xldot = xni + d.xfact;
        xnddt = xnddt * xldot;
    };
    while (fabs(t - atime) >= kStepp) {
        dots();
        xli = xli + xldot * delt + xndot * kStep2;
        xni = xni + xndot * delt + xnddt * kStep2;
        atime += delt;
    }
    const double ft = t - atime;
    dots();
    *xn = xni + xndot * ft + xnddt * ft * ft * 0.5;
    const double xl = xli + xldot * ft + xndot * ft * ft * 0.5;
    const double temp = -*xnodes + d.thgr + t * kThdt;
    *xll = d.isynfl ? xl - *omgasm + temp : xl + temp + temp;
}

// DEEP entry DPPER: lunar-solar periodics.  (The report re-evaluates them only when t has moved 30 min since the last
// evaluation -- a saving on 1980 hardware that makes results depend on the call history; they are evaluated every time here.)
void Sgp4::deep_periodic(double t, double* em, double* xinc, double* omgasm, double* xnodes, double* xll) const
{
    const DeepSpace& d = ds_;
    const double sinis = sin(*xinc), cosis = cos(*xinc);
    double zm = d.zmos + kZns * t;
    double zf = zm + 2.0 * kZes * sin(zm);
    double sinzf = sin(zf);
    double f2 = 0.5 * sinzf * sinzf - 0.25;
    double f3 = -0.5 * sinzf * cos(zf);
    const double ses = d.se2 * f2 + d.se3 * f3;
    const double sis = d.si2 * f2 + d.si3 * f3;
    const double sls = d.sl2 * f2 + d.sl3 * f3 + d.sl4 * sinzf;
    const double sghs = d.sgh2 * f2 + d.sgh3 * f3 + d.sgh4 * sinzf;
    const double shs = d.sh2 * f2 + d.sh3 * f3;
    zm = d.zmol + kZnl * t;
    zf = zm + 2.0 * kZel * sin(zm);
    sinzf = sin(zf);
    f2 = 0.5 * sinzf * sinzf - 0.25;
    f3 = -0.5 * sinzf * cos(zf);
    const double sel = d.ee2 * f2 + d.e3 * f3;
    const double sil = d.xi2 * f2 + d.xi3 * f3;
    const double sll = d.xl2 * f2 + d.xl3 * f3 + d.xl4 * sinzf;
    const double sghl = d.xgh2 * f2 + d.xgh3 * f3 + d.xgh4 * sinzf;
    const double shl = d.xh2 * f2 + d.xh3 * f3;
    const double pe = ses + sel;
    const double pinc = sis + sil;
    const double pl = sls + sll;
    double pgh = sghs + sghl;
    double ph = shs + shl;
    *xinc += pinc;
    *em += pe;
    if (d.xqncl >= 0.2) {
        // apply periodics directly
        ph = ph / d.siniq;
        pgh = pgh - d.cosiq * ph;
        *omgasm += pgh;
        *xnodes += ph;
        *xll += pl;
    } else {
        // apply periodics with the Lyddane modification (low inclinations)
        const double sinok = sin(*xnodes), cosok = cos(*xnodes);
        double alfdp = sinis * sinok;
        double betdp = sinis * cosok;
        const double dalf = ph * cosok + pinc * cosis * sinok;
        const double dbet = -ph * sinok + pinc * cosis * cosok;
        alfdp += dalf;
        betdp += dbet;
        double xls = *xll + *omgasm + cosis * *xnodes;
        const double dls = pl + pgh - pinc * *xnodes * sinis;
        xls += dls;
        *xnodes = actan(alfdp, betdp);
        *xll += pl;
        *omgasm = xls - *xll - cos(*xinc) * *xnodes;
    }
}

Observation observe(const Vec3& pos, const Vec3& vel, double jd, double lat_deg, double lon_deg, double alt_m, const Constants& k)
{
    const double kXkmper = k.xkmper, kF = k.f;
    const double lat = lat_deg * kDeg, lon = lon_deg * kDeg, alt_km = alt_m / 1000.0;
    // observer position and velocity in the same inertial frame
    const double theta = fmod2p(theta_g(jd) + lon);   // local mean sidereal time
    const double sinlat = sin(lat), coslat = cos(lat);
    const double c = 1.0 / sqrt(1.0 + kF * (kF - 2.0) * sinlat * sinlat);
    const double sq = (1.0 - kF) * (1.0 - kF) * c;
    const double achcp = (kXkmper * c + alt_km) * coslat;
    const Vec3 opos{achcp * cos(theta), achcp * sin(theta), (kXkmper * sq + alt_km) * sinlat};
    const double mfactor = k.mfactor;   // rad / s
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
    return dorbit::observe(p, v, jd, lat_, lon_, alt_, sgp4_.consts());
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

int doppler_b200_orbit_constants(int which)
{
    const int prev = dorbit::default_constant_set();
    if (which == DOPPLER_B200_ORBIT_GPREDICT || which == DOPPLER_B200_ORBIT_WGS72) dorbit::set_default_constant_set(which);
    return prev;
}

int doppler_b200_tracker_is_deep_space(const doppler_b200_tracker* tr) { return tr ? (int)tr->t.deep_space() : -1; }

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
