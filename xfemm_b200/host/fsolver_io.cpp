// FSolver file I/O: .fem problem description, fmesher mesh files, .ans writers.
// Behavioural mirror of (all host side, no arithmetic beyond unit conversion):
//   FEASolver::LoadProblemFile  cfemm/libfemm/feasolver.cpp:181-563 (+ the C*Prop::fromStream parsers)
//   FSolver::LoadProblemFile    cfemm/fsolver/fsolver.cpp:202-348
//   FSolver::LoadMesh           cfemm/fsolver/fsolver.cpp:350-718
//   FSolver::WriteStatic2D      cfemm/fsolver/static2d.cpp:1038-1195
//   FSolver::WriteHarmonic2D    cfemm/fsolver/harmonic2d.cpp:892-1036
// The parser here is line oriented ("key = value" lines inside [section] / <Begin…>…<End…>
// blocks) instead of the reference's token stream; it accepts the same files.
#include "fsolver.h"
#include <chrono>
#include <unordered_map>
#include "../csrc/layout.h"
#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace femm
{
const double LengthConvMeters[6] = {0.0254, 0.001, 0.01, 1.0, 2.54e-5, 1.e-6};
int PrintWarningMsg(const char *message, ...) { return printf("%s", message); }
} // namespace femm

using namespace femm;
using namespace femmsolver;

namespace
{
std::string trimmed(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
std::string lowered(std::string s)
{
    for (auto &c : s) c = (char)tolower((unsigned char)c);
    return s;
}
// splits "  <Key> = value" into lower-cased key and trimmed value; key only if there is no '='
void split_kv(const std::string &line, std::string &key, std::string &val)
{
    size_t eq = line.find('=');
    if (eq == std::string::npos) { key = lowered(trimmed(line)); val.clear(); return; }
    key = lowered(trimmed(line.substr(0, eq)));
    val = trimmed(line.substr(eq + 1));
}
bool to_double(const std::string &v, double &out)
{
    try { out = std::stod(v); return true; } catch (...) { return false; }
}
bool to_int(const std::string &v, int &out)
{
    try { out = std::stoi(v); return true; } catch (...) { return false; }
}
std::string unquote(const std::string &v)
{
    size_t a = v.find('"');
    size_t b = v.rfind('"');
    if (a == std::string::npos || b == a) return std::string();
    return v.substr(a + 1, b - a - 1);
}

struct LineReader
{
    std::ifstream in;
    bool next(std::string &line)
    {
        while (std::getline(in, line))
        {
            line = trimmed(line);
            if (!line.empty()) return true;
        }
        return false;
    }
};

// reads "<key> = value" lines until `endtag`; handler returns false for unknown keys
template <typename F>
bool read_block(LineReader &R, const char *begintag, const char *endtag, std::ostringstream &err, F handler)
{
    std::string line, key, val;
    if (!R.next(line)) return false;
    if (lowered(line) != begintag) { err << "Expected token " << begintag << ", but got " << line << "\n"; return false; }
    while (R.next(line))
    {
        split_kv(line, key, val);
        if (key == endtag) return true;
        if (!handler(key, val)) err << "unexpected token: " << key << "\n";
    }
    return false;
}
} // namespace

double Power(double x, int n) { return pow(x, (double)n); }

FSolver::FSolver()
{
    CleanUp();
    Frequency = 0.0;
    Relax = 0.0;
    NumCircPropsOrig = 0;
    WarnMessage = &PrintWarningMsg;
    PrintMessage = &PrintWarningMsg;
    precondOption = -1;
    maxIterOption = 0;
    // one mesh over several GPUs (see fsolver.h): the command line and unmodified callers pick it up from the environment
    if (const char *w = getenv("XFEMM_DIST_WORLD")) distWorld = std::max(1, atoi(w));
    if (const char *r = getenv("XFEMM_DIST_RANK")) distRank = std::max(0, atoi(r));
    magdev = nullptr;
    cmagdev = nullptr;
    keepSystem = false;
    ageUnionPattern = false;
    assembleOnly = false;
    localityOrder = -1;
    fetchSolution = true;
    cacheStaticPart = true;
    warmResolve = true;
    resolveStart = false;
    extrapolateResolve = true;
}

FSolver::~FSolver() { releaseDevice(); }

void FSolver::CleanUp()
{
    FileFormat = -1; Precision = 1.e-08; MinAngle = 0.; Depth = -1;
    LengthUnits = LengthInches; Coords = CART; ProblemType = PLANAR;
    extZo = extRo = extRi = 0.0;
    comment.clear();
    ACSolver = 0;
    DoForceMaxMeshArea = false; DoSmartMesh = true; bMultiplyDefinedLabels = false;
    BandWidth = 0;
    meshele.clear();
    NumNodes = NumEls = 0;
    NumBlockProps = NumPBCs = NumLineProps = NumPointProps = NumCircProps = NumBlockLabels = 0;
    NumAirGapElems = 0;
    pbclist.clear(); agelist.clear();
    PrevType = 0;
    nodeproplist.clear(); lineproplist.clear(); blockproplist.clear(); circproplist.clear(); labellist.clear();
}

std::string FSolver::getErrorString(LoadMeshErr err)
{
    switch (err)
    {
    case NOERROR: return std::string();
    case BADEDGEFILE: return "problem loading mesh:\nCould not open .edge file.\n";
    case BADELEMENTFILE: return "problem loading mesh:\nCould not open .ele file.\n";
    case BADFEMFILE: return "problem loading mesh:\nCould not open .fem file.\n";
    case BADNODEFILE: return "problem loading mesh:\nCould not open .node file.\n";
    case BADPBCFILE: return "problem loading mesh:\nCould not open .pbc file.\n";
    case MISSINGMATPROPS: return "problem loading mesh:\nMaterial properties have not been defined for all regions.\n";
    case ELMLABELTOOBIG: return "problem loading mesh:\nAn element had a label index exceeding the number of labels.\n";
    }
    return "unknown error";
}

bool FSolver::parseFem(const std::string &file)
{
    LineReader R;
    R.in.open(file.c_str());
    std::ostringstream err;
    if (!R.in.is_open())
    {
        err << "Couldn't read from specified .fem file: " << file << "\n";
        WarnMessage(err.str().c_str());
        return false;
    }
    CleanUp();
    Frequency = 0.0;
    std::string line, key, val;
    bool ok = true;
    auto need_d = [&](double &dst) { if (!to_double(val, dst)) { err << "Could not convert '" << val << "' to double\n"; ok = false; } };
    auto need_i = [&](int &dst) { if (!to_int(val, dst)) { err << "Could not convert '" << val << "' to int\n"; ok = false; } };
    while (ok && R.next(line))
    {
        split_kv(line, key, val);
        if (key == "[format]") { need_d(FileFormat); continue; }
        if (key == "[precision]") { need_d(Precision); continue; }
        if (key == "[minangle]") { need_d(MinAngle); continue; }
        if (key == "[depth]") { need_d(Depth); continue; }
        if (key == "[frequency]") { need_d(Frequency); continue; } // fsolver.cpp:1355-1366
        if (key == "[lengthunits]")
        {
            std::string u = lowered(val);
            if (u == "inches") LengthUnits = LengthInches;
            else if (u == "millimeters") LengthUnits = LengthMillimeters;
            else if (u == "centimeters") LengthUnits = LengthCentimeters;
            else if (u == "mils") LengthUnits = LengthMils;
            else if (u == "microns") LengthUnits = LengthMicrometers;
            else if (u == "meters") LengthUnits = LengthMeters;
            continue;
        }
        if (key == "[coordinates]")
        {
            std::string u = lowered(val);
            if (u == "cartesian") Coords = CART;
            if (u == "polar") Coords = POLAR;
            continue;
        }
        if (key == "[problemtype]")
        {
            std::string u = lowered(val);
            if (u == "planar") ProblemType = PLANAR;
            if (u == "axisymmetric") ProblemType = AXISYMMETRIC;
            continue;
        }
        if (key == "[extzo]") { need_d(extZo); continue; }
        if (key == "[extro]") { need_d(extRo); continue; }
        if (key == "[extri]") { need_d(extRi); continue; }
        if (key == "[comment]") { comment = unquote(val); continue; }
        if (key == "[acsolver]") { need_i(ACSolver); continue; }
        if (key == "[prevtype]") { need_i(PrevType); continue; }
        if (key == "[prevsoln]") { previousSolutionFile = unquote(val); continue; }
        if (key == "[forcemaxmesh]") { int v = 0; need_i(v); DoForceMaxMeshArea = v != 0; continue; }
        if (key == "[dosmartmesh]") { int v = 0; need_i(v); DoSmartMesh = v != 0; continue; }
        if (key == "[pointprops]")
        {
            int k = 0; need_i(k);
            for (int q = 0; ok && q < k; q++)
            {
                CMPointProp P;
                ok = read_block(R, "<beginpoint>", "<endpoint>", err, [&](const std::string &kk, const std::string &vv) {
                    if (kk == "<pointname>") { P.PointName = unquote(vv); return true; }
                    if (kk == "<a_re>") return to_double(vv, P.A.re);
                    if (kk == "<a_im>") return to_double(vv, P.A.im);
                    if (kk == "<i_re>") return to_double(vv, P.J.re);
                    if (kk == "<i_im>") return to_double(vv, P.J.im);
                    return false;
                });
                if (ok) { nodeproplist.push_back(P); NumPointProps++; }
            }
            if (NumPointProps != k) { err << "Expected " << k << " PointProps, but got " << NumPointProps << "\n"; ok = false; }
            continue;
        }
        if (key == "[bdryprops]")
        {
            int k = 0; need_i(k);
            for (int q = 0; ok && q < k; q++)
            {
                CMBoundaryProp P;
                ok = read_block(R, "<beginbdry>", "<endbdry>", err, [&](const std::string &kk, const std::string &vv) {
                    if (kk == "<bdrytype>") return to_int(vv, P.BdryFormat);
                    if (kk == "<mu_ssd>") return to_double(vv, P.Mu);
                    if (kk == "<sigma_ssd>") return to_double(vv, P.Sig);
                    if (kk == "<a_0>") return to_double(vv, P.A0);
                    if (kk == "<a_1>") return to_double(vv, P.A1);
                    if (kk == "<a_2>") return to_double(vv, P.A2);
                    if (kk == "<phi>") return to_double(vv, P.phi);
                    if (kk == "<c0>") return to_double(vv, P.c0.re);
                    if (kk == "<c1>") return to_double(vv, P.c1.re);
                    if (kk == "<c0i>") return to_double(vv, P.c0.im);
                    if (kk == "<c1i>") return to_double(vv, P.c1.im);
                    if (kk == "<innerangle>") return to_double(vv, P.InnerAngle);
                    if (kk == "<outerangle>") return to_double(vv, P.OuterAngle);
                    if (kk == "<bdryname>") { P.BdryName = unquote(vv); return true; }
                    return false;
                });
                if (ok) { lineproplist.push_back(P); NumLineProps++; }
            }
            if (NumLineProps != k) { err << "Expected " << k << " BoundaryProps, but got " << NumLineProps << "\n"; ok = false; }
            continue;
        }
        if (key == "[blockprops]")
        {
            int k = 0; need_i(k);
            for (int q = 0; ok && q < k; q++)
            {
                CMSolverMaterialProp P;
                ok = read_block(R, "<beginblock>", "<endblock>", err, [&](const std::string &kk, const std::string &vv) {
                    if (kk == "<blockname>") { P.BlockName = unquote(vv); return true; }
                    if (kk == "<mu_x>") return to_double(vv, P.mu_x);
                    if (kk == "<mu_y>") return to_double(vv, P.mu_y);
                    if (kk == "<h_c>") return to_double(vv, P.H_c);
                    if (kk == "<h_cangle>") return to_double(vv, P.Theta_m);
                    if (kk == "<j_re>") return to_double(vv, P.J.re);
                    if (kk == "<j_im>") return to_double(vv, P.J.im);
                    if (kk == "<sigma>") return to_double(vv, P.Cduct);
                    if (kk == "<phi_h>") return to_double(vv, P.Theta_hn);
                    if (kk == "<phi_hx>") return to_double(vv, P.Theta_hx);
                    if (kk == "<phi_hy>") return to_double(vv, P.Theta_hy);
                    if (kk == "<d_lam>") return to_double(vv, P.Lam_d);
                    if (kk == "<lamfill>") return to_double(vv, P.LamFill);
                    if (kk == "<wired>") return to_double(vv, P.WireD);
                    if (kk == "<lamtype>") return to_int(vv, P.LamType);
                    if (kk == "<nstrands>") return to_int(vv, P.NStrands);
                    if (kk == "<bhpoints>")
                    {
                        if (!to_int(vv, P.BHpoints)) return false;
                        // the B-H table follows as whitespace separated number pairs
                        for (int i = 0; i < P.BHpoints; i++)
                        {
                            double b = 0, h = 0;
                            R.in >> b >> h;
                            P.Bdata.push_back(b);
                            P.Hdata.push_back(CComplex(h));
                        }
                        return true;
                    }
                    return false;
                });
                if (ok) { blockproplist.push_back(P); NumBlockProps++; }
            }
            if (NumBlockProps != k) { err << "Expected " << k << " BlockProps, but got " << NumBlockProps << "\n"; ok = false; }
            continue;
        }
        if (key == "[circuitprops]" || key == "[conductorprops]")
        {
            int k = 0; need_i(k);
            for (int q = 0; ok && q < k; q++)
            {
                CMCircuit P;
                ok = read_block(R, "<begincircuit>", "<endcircuit>", err, [&](const std::string &kk, const std::string &vv) {
                    if (kk == "<circuitname>") { P.CircName = unquote(vv); return true; }
                    if (kk == "<voltgradient_re>") return to_double(vv, P.dVolts.re);
                    if (kk == "<voltgradient_im>") return to_double(vv, P.dVolts.im);
                    if (kk == "<totalamps_re>") return to_double(vv, P.Amps.re);
                    if (kk == "<totalamps_im>") return to_double(vv, P.Amps.im);
                    if (kk == "<circuittype>") return to_int(vv, P.CircType);
                    return false;
                });
                if (ok) { circproplist.push_back(P); NumCircProps++; }
            }
            if (NumCircProps != k) { err << "Expected " << k << " CircuitProps, but got " << NumCircProps << "\n"; ok = false; }
            continue;
        }
        if (key == "[numblocklabels]")
        {
            need_i(NumBlockLabels);
            for (int q = 0; ok && q < NumBlockLabels; q++)
            {
                // CBlockLabel.cpp:111-155: x y blocktype maxarea circuit magdir group turns extDefault ["fctn"]
                std::string l;
                if (!R.next(l)) { err << "Expected " << NumBlockLabels << " BlockLabels, but got " << q << "\n"; ok = false; break; }
                std::istringstream ls(l);
                CMBlockLabel B;
                ls >> B.x >> B.y >> B.BlockType;
                B.BlockType--;
                ls >> B.MaxArea;
                if (B.MaxArea <= 0) B.MaxArea = 0; else B.MaxArea *= PI * B.MaxArea / 4.;
                ls >> B.InCircuit;
                B.InCircuit--;
                ls >> B.MagDir >> B.InGroup >> B.Turns;
                int extDefault = 0;
                ls >> extDefault;
                B.IsDefault = (extDefault & 2) != 0;
                B.IsExternal = (extDefault & 1) != 0;
                std::string rest;
                std::getline(ls, rest);
                B.MagDirFctn = unquote(rest);
                labellist.push_back(B);
            }
            continue;
        }
        if (key == "[numpoints]" || key == "[numsegments]" || key == "[numarcsegments]" || key == "[numholes]")
        {
            int k = 0; need_i(k);
            std::string skip;
            while (0 < k--) std::getline(R.in, skip);
            continue;
        }
        err << "Unknown token: " << key << "\nContext line:\n" << line << "\n";
        ok = false;
    }
    if (!ok)
    {
        std::string msg = "Parse error while reading input file " + file + "!\n" + err.str();
        WarnMessage(msg.c_str());
        return false;
    }
    if (!err.str().empty()) WarnMessage(err.str().c_str());
    return true;
}

bool FSolver::LoadProblemFile()
{
    Relax = 1.;
    std::string femFile = PathName + ".fem";
    if (!parseFem(femFile)) return false;
    meshLoadedFromPrevSolution = false;
    if (!previousSolutionFile.empty())
    {
        // [PrevType] 1 / 2 (incremental / frozen permeability) do not run in the reference itself (DESIGN.md §2, row b)
        if (PrevType != 0)
        {
            WarnMessage("xfemm_b200: [PrevSoln] with [PrevType] 1 or 2 (incremental / frozen permeability) is not supported\n");
            return false;
        }
        // mesh reuse.  The reference returns from here without the B-H set-up and the circuit expansion below
        // (fsolver.cpp:237), which leaves a problem with B-H materials or series circuits unusable there; both are
        // done here as for any other problem.
        if (!loadPreviousSolution(false)) return false;
    }
    // B-H spline set-up on the host (fsolver.cpp:243-276)
    for (auto &prop : blockproplist)
        if (prop.BHpoints > 0)
        {
            prop.GetSlopes(Frequency * 2. * PI);
            prop.MuMax = 0;
        }
    if (NumCircProps == 0) return true;

    // Series circuits become one parallel "circuit" per block label (fsolver.cpp:278-317).
    circproplist.resize(NumCircProps + NumBlockLabels);
    for (int k = 0; k < NumCircProps; k++) circproplist[k].OrigCirc = -1;
    NumCircPropsOrig = NumCircProps;
    for (int k = 0; k < NumBlockLabels; k++)
        if (labellist[k].InCircuit >= 0)
        {
            int ic = labellist[k].InCircuit;
            if (ic < (int)circproplist.size() && circproplist[ic].CircType == 1)
            {
                CMCircuit nc = circproplist[ic];
                nc.OrigCirc = ic;
                nc.Amps.im *= labellist[k].Turns;
                nc.Amps.re *= labellist[k].Turns;
                circproplist[NumCircProps] = nc;
                labellist[k].InCircuit = NumCircProps;
                NumCircProps++;
            }
        }
    for (int k = 0; k < NumCircProps; k++)
        if (circproplist[k].CircType == 1) circproplist[k].CircType = 0;
    return true;
}

// ---------------------------------------------------------------------------------------
// mesh files.  A small hand-rolled number scanner replaces fscanf: at 2 M nodes the text
// parse is a visible part of the end-to-end time.
// ---------------------------------------------------------------------------------------
namespace
{
struct TextFile
{
    std::vector<char> buf;
    const char *p = nullptr, *end = nullptr;
    bool open(const std::string &name)
    {
        FILE *f = fopen(name.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize((size_t)sz + 1);
        size_t got = fread(buf.data(), 1, (size_t)sz, f);
        fclose(f);
        buf[got] = 0;
        p = buf.data();
        end = p + got;
        return true;
    }
    void skipws() { while (p < end && isspace((unsigned char)*p)) p++; }
    bool nextInt(int &v)
    {
        skipws();
        if (p >= end) return false;
        char *q;
        long r = strtol(p, &q, 10);
        if (q == p) return false;
        // tolerate "12.0"-style tokens the way %i would stop at the dot
        p = q;
        v = (int)r;
        return true;
    }
    bool nextDouble(double &v)
    {
        skipws();
        if (p >= end) return false;
        char *q;
        v = strtod(p, &q);
        if (q == p) return false;
        p = q;
        return true;
    }
    std::string restOfLine()
    {
        const char *s = p;
        while (p < end && *p != '\n') p++;
        std::string r(s, p);
        if (p < end) p++;
        return r;
    }
    void skipLine() { restOfLine(); }
};

// ---- line-parallel parsing of the big mesh files -------------------------------------------
// The reference reads .node/.ele/.edge with fgets + sscanf, one record per line (fsolver.cpp:362-697).
// Here the line starts are located once and the records are parsed by a handful of host threads
// (the files are 100-150 MB of text at 2 M nodes; a serial scan was 1.5 s, most of "load mesh").
static std::vector<const char *> line_starts(const char *b, const char *e)
{
    const size_t len = (size_t)(e - b);
    const int chunks = 64;
    std::vector<size_t> cnt(chunks + 1, 0);
    xb::parallel_for(chunks, [&](int c0, int c1, int) {
        for (int c = c0; c < c1; c++)
        {
            const char *p = b + len * c / chunks, *q = b + len * (c + 1) / chunks;
            size_t n = 0;
            while (p < q && (p = (const char *)memchr(p, '\n', (size_t)(q - p))) != nullptr) { n++; p++; }
            cnt[c + 1] = n;
        }
    }, 1);
    for (int c = 0; c < chunks; c++) cnt[c + 1] += cnt[c];
    std::vector<const char *> ls(cnt[chunks] + 2);
    ls[0] = b;
    xb::parallel_for(chunks, [&](int c0, int c1, int) {
        for (int c = c0; c < c1; c++)
        {
            const char *p = b + len * c / chunks, *q = b + len * (c + 1) / chunks;
            size_t k = cnt[c] + 1;
            while (p < q && (p = (const char *)memchr(p, '\n', (size_t)(q - p))) != nullptr) { p++; ls[k++] = p; }
        }
    }, 1);
    ls[cnt[chunks] + 1] = e;
    ls.resize(cnt[chunks] + 1); // ls[i] = start of line i (a last entry after the final newline may be empty)
    return ls;
}
static inline const char *skip_blank(const char *p) { while (*p == ' ' || *p == '\t' || *p == '\r') p++; return p; }
static inline bool parse_int(const char *&p, int &v)
{
    p = skip_blank(p);
    bool neg = false;
    if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
    if (*p < '0' || *p > '9') return false;
    long r = 0;
    while (*p >= '0' && *p <= '9') r = r * 10 + (*p++ - '0');
    v = (int)(neg ? -r : r);
    return true;
}
static inline bool parse_double(const char *&p, double &v)
{
    p = skip_blank(p);
    char *q;
    v = strtod(p, &q);
    if (q == p) return false;
    p = q;
    return true;
}
} // namespace

// Mesh out of the [Solution] section of an earlier .ans file (reference: FSolver::loadPreviousSolution and the four
// Load...FromSolution routines, fsolver.cpp:720-1081).  Line-oriented like the reference's fgets/sscanf: a static
// .ans has 4 numbers per element line, a harmonic one 8; what a line does not hold keeps the element's default —
// e[] = {0,0,0}, i.e. boundary property 0 on every side, exactly as the reference's CElement() leaves it.
bool FSolver::loadPreviousSolution(bool loadAprev)
{
    if (previousSolutionFile.empty()) return false;
    FILE *fp = fopen(previousSolutionFile.c_str(), "rt");
    if (!fp)
    {
        WarnMessage(("Failed to open the specified previous solution file, file path was:\n" + previousSolutionFile + "\n").c_str());
        return false;
    }
    char s[1024], q[256];
    bool hasSolution = false;
    while (fgets(s, sizeof s, fp))
    {
        q[0] = 0;
        sscanf(s, "%255s", q);
        std::string key(q);
        for (auto &ch : key) ch = (char)tolower((unsigned char)ch);
        if (key.compare(0, 11, "[frequency]") == 0)
        {
            double prevFreq = 0;
            const char *eq = strchr(s, '=');
            if (eq) sscanf(eq + 1, "%lf", &prevFreq);
            if (prevFreq != 0)
            {
                fclose(fp);
                WarnMessage(("Previous solution file (" + previousSolutionFile + ") appears to be an AC problem, only DC previous solutions are presently supported\n").c_str());
                return false;
            }
        }
        if (key.compare(0, 10, "[solution]") == 0) { hasSolution = true; break; }
    }
    if (!hasSolution)
    {
        fclose(fp);
        WarnMessage(("No solution was found in previous solution file, file path was:\n" + previousSolutionFile + "\n").c_str());
        return false;
    }
    auto line = [&]() { return fgets(s, sizeof s, fp) != nullptr; };
    const double toCm = 100. * LengthConvMeters[LengthUnits];
    // nodes: x y A marker
    NumNodes = 0;
    if (line()) sscanf(s, "%i", &NumNodes);
    Aprev.clear();
    meshnode.assign((size_t)std::max(0, NumNodes), CNode());
    for (int i = 0; i < NumNodes && line(); i++)
    {
        double a = 0;
        CNode &nd = meshnode[i];
        sscanf(s, "%lf %lf %lf %i", &nd.x, &nd.y, &a, &nd.BoundaryMarker);
        nd.x *= toCm; nd.y *= toCm;
        if (loadAprev) Aprev.push_back(a);
    }
    // elements
    NumEls = 0;
    if (line()) sscanf(s, "%i", &NumEls);
    meshele.assign((size_t)std::max(0, NumEls), femmsolver::CMElement());
    for (int i = 0; i < NumEls && line(); i++)
    {
        femmsolver::CMElement &e = meshele[i];
        e.e[0] = e.e[1] = e.e[2] = 0; // the reference's constructor default, kept when the line has 4 fields only
        sscanf(s, "%i %i %i %i %i %i %i %lf", &e.p[0], &e.p[1], &e.p[2], &e.lbl, &e.e[0], &e.e[1], &e.e[2], &e.Jprev);
        if (e.lbl < 0 || e.lbl >= (int)labellist.size()) { fclose(fp); WarnMessage("previous solution: element with an unknown block label\n"); return false; }
        e.blk = labellist[e.lbl].BlockType;
    }
    // per-label circuit results: skipped
    int numLabels = 0;
    if (line()) sscanf(s, "%i", &numLabels);
    for (int i = 0; i < numLabels; i++) line();
    // periodic pairs
    NumPBCs = 0;
    pbclist.clear();
    if (line())
    {
        sscanf(s, "%i", &NumPBCs);
        for (int i = 0; i < NumPBCs && line(); i++)
        {
            CCommonPoint c;
            sscanf(s, "%i %i %i", &c.x, &c.y, &c.t);
            pbclist.push_back(c);
        }
    }
    // air-gap elements
    NumAirGapElems = 0;
    agelist.clear();
    if (line()) sscanf(s, "%i", &NumAirGapElems);
    for (int i = 0; i < NumAirGapElems; i++)
    {
        femmsolver::CAirGapElement age;
        if (!line()) break;
        age.BdryName = s;
        if (!line()) break;
        sscanf(s, "%i %lf %lf %lf %lf %lf %lf %lf %i %lf %lf", &age.BdryFormat, &age.InnerAngle, &age.OuterAngle, &age.ri, &age.ro,
               &age.totalArcLength, &age.agc.re, &age.agc.im, &age.totalArcElements, &age.InnerShift, &age.OuterShift);
        for (int k = 0; k <= age.totalArcElements && line(); k++)
        {
            CQuadPoint qp = {0, 0, 0, 0, 0, 0, 0, 0};
            sscanf(s, "%i %lf %i %lf %i %lf %i %lf", &qp.n0, &qp.w0, &qp.n1, &qp.w1, &qp.n2, &qp.w2, &qp.n3, &qp.w3);
            if (qp.n0 < 0 || qp.n1 < 0 || qp.n2 < 0 || qp.n3 < 0)
            {
                fclose(fp);
                WarnMessage("An error occured while reading pbc file, quadNode has negative node number.\n");
                return false;
            }
            age.quadNode.push_back(qp);
        }
        agelist.push_back(age);
    }
    fclose(fp);
    // the file is in solver numbering already: no renumbering, and no bandwidth was computed (the reference creates
    // its linear problem with the constructor's BandWidth = 0, which its SetValue reads as "scan every row")
    BandWidth = 0;
    cmPermutation.resize((size_t)NumNodes);
    for (int i = 0; i < NumNodes; i++) cmPermutation[i] = i;
    edgeList.clear();
    meshLoadedFromPrevSolution = true;
    return true;
}

LoadMeshErr FSolver::LoadMesh(bool deleteFiles)
{
    if (meshLoadedFromPrevSolution) return NOERROR; // fsolver.cpp:357
    TextFile F;
    int k = 0, j = 0;
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tt = now();
    auto tick = [&](const char *what) { if (timing) { fprintf(stderr, "[xfemm timing]   LoadMesh %s %.3f s\n", what, now() - tt); tt = now(); } };
    // ---- nodes (fsolver.cpp:362-391): coordinates to centimetres
    if (!F.open(PathName + ".node")) return BADNODEFILE;
    F.nextInt(k);
    F.skipLine();
    NumNodes = k;
    meshnode.assign(k, CNode());
    const double cm = 100 * LengthConvMeters[LengthUnits];
    {
        std::vector<const char *> ls = line_starts(F.buf.data(), F.end);
        if ((int)ls.size() < k + 1) return BADNODEFILE;
        xb::parallel_for(k, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                CNode &nd = meshnode[i];
                const char *p = ls[i + 1];
                int idx = 0, mk = 0;
                parse_int(p, idx);
                parse_double(p, nd.x);
                parse_double(p, nd.y);
                parse_int(p, mk);
                nd.BoundaryMarker = (mk > 1) ? mk - 2 : -1;
                nd.x *= cm;
                nd.y *= cm;
            }
        });
    }
    tick("nodes");
    // ---- periodic pairs + air-gap elements (fsolver.cpp:393-515)
    if (!F.open(PathName + ".pbc")) return BADPBCFILE;
    F.nextInt(NumPBCs);
    F.skipLine();
    pbclist.clear();
    for (int i = 0; i < NumPBCs; i++)
    {
        CCommonPoint pt;
        F.nextInt(j); F.nextInt(pt.x); F.nextInt(pt.y); F.nextInt(pt.t);
        F.skipLine();
        pbclist.push_back(pt);
    }
    NumAirGapElems = 0;
    F.nextInt(NumAirGapElems);
    F.skipLine();
    agelist.clear();
    for (int i = 0; i < NumAirGapElems; i++)
    {
        CAirGapElement age;
        age.BdryName = F.restOfLine() + "\n"; // the reference keeps the line incl. its newline
        double fmt = 0, nel = 0;
        F.nextDouble(fmt);
        age.BdryFormat = (int)fmt;
        F.nextDouble(age.InnerAngle); F.nextDouble(age.OuterAngle); F.nextDouble(age.ri); F.nextDouble(age.ro);
        F.nextDouble(age.totalArcLength); F.nextDouble(age.agc.re); F.nextDouble(age.agc.im);
        F.nextDouble(nel);
        age.totalArcElements = (int)nel;
        F.nextDouble(age.InnerShift); F.nextDouble(age.OuterShift);
        F.skipLine();
        for (int q = 0; q <= age.totalArcElements; q++)
        {
            CQuadPoint qp;
            F.nextInt(qp.n0); F.nextDouble(qp.w0); F.nextInt(qp.n1); F.nextDouble(qp.w1);
            F.nextInt(qp.n2); F.nextDouble(qp.w2); F.nextInt(qp.n3); F.nextDouble(qp.w3);
            F.skipLine();
            if (qp.n0 < 0 || qp.n1 < 0 || qp.n2 < 0 || qp.n3 < 0)
            {
                WarnMessage("An error occured while reading file, quadNode has negative node number.\n");
                return BADPBCFILE;
            }
            age.quadNode.push_back(qp);
        }
        agelist.push_back(age);
    }
    // ---- elements (fsolver.cpp:517-622)
    if (!F.open(PathName + ".ele")) return BADELEMENTFILE;
    F.nextInt(k);
    F.skipLine();
    NumEls = k;
    meshele.assign(k, CMElement());
    int defaultLabel = -1;
    for (int i = 0; i < NumBlockLabels; i++)
        if (labellist[i].IsDefault) defaultLabel = i;
    auto removeMeshFiles = [&]() {
        for (const char *ext : {".ele", ".node", ".pbc", ".poly", ".edge"}) remove((PathName + ext).c_str());
    };
    {
        std::vector<const char *> ls = line_starts(F.buf.data(), F.end);
        if ((int)ls.size() < k + 1) return BADELEMENTFILE;
        xb::parallel_for(k, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                CMElement &el = meshele[i];
                const char *p = ls[i + 1];
                int idx = 0;
                parse_int(p, idx);
                parse_int(p, el.p[0]); parse_int(p, el.p[1]); parse_int(p, el.p[2]);
                parse_int(p, el.lbl);
            }
        });
    }
    for (int i = 0; i < k; i++)
    {
        CMElement &el = meshele[i];
        el.lbl--;
        if (el.lbl < 0) el.lbl = defaultLabel;
        if (el.lbl < 0)
        {
            char buf[256];
            snprintf(buf, sizeof buf, "Material properties have not been defined for all regions.\nThe element number %i had label %i\n", i, el.lbl);
            WarnMessage(buf);
            if (deleteFiles) removeMeshFiles();
            return MISSINGMATPROPS;
        }
        if (!(el.lbl < (int)labellist.size()))
        {
            char buf[256];
            snprintf(buf, sizeof buf, "The element number %i had label %i which is greater than the number of available labels (%i)\n",
                     i + 1, el.lbl + 1, (int)labellist.size());
            WarnMessage(buf);
            if (deleteFiles) removeMeshFiles();
            return ELMLABELTOOBIG;
        }
        el.blk = labellist[el.lbl].BlockType;
        el.e[0] = el.e[1] = el.e[2] = -1;
        el.mu1 = -1.; el.mu2 = -1.;
    }
    tick("elements");
    // ---- boundary edges (fsolver.cpp:624-697).  The reference walks the elements around the first node
    // of every tagged edge; equivalently every element side looks its node pair up among the tagged
    // edges (a few thousand of millions; a later edge of the file overrides an earlier one).
    if (!F.open(PathName + ".edge")) return BADEDGEFILE;
    F.nextInt(k);
    F.nextInt(j);
    edgeList.assign((size_t)k * 2, 0);
    std::vector<int> edgeMark(k, 0);
    {
        std::vector<const char *> ls = line_starts(F.buf.data(), F.end);
        if ((int)ls.size() < k + 1) return BADEDGEFILE;
        xb::parallel_for(k, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                const char *p = ls[i + 1];
                int idx = 0;
                parse_int(p, idx);
                parse_int(p, edgeList[2 * (size_t)i]); parse_int(p, edgeList[2 * (size_t)i + 1]);
                parse_int(p, edgeMark[i]);
            }
        });
    }
    {
        std::unordered_map<long long, int> tagged;
        for (int i = 0; i < k; i++)
            if (edgeMark[i] < 0)
            {
                const long long a = std::min(edgeList[2 * (size_t)i], edgeList[2 * (size_t)i + 1]);
                const long long b = std::max(edgeList[2 * (size_t)i], edgeList[2 * (size_t)i + 1]);
                tagged[a * (long long)NumNodes + b] = -(edgeMark[i] + 2);
            }
        if (!tagged.empty())
        {
            std::vector<char> onTagged(NumNodes, 0); // cheap pre-filter: nodes that sit on a tagged edge
            for (const auto &kv : tagged) { onTagged[kv.first / NumNodes] = 1; onTagged[kv.first % NumNodes] = 1; }
            xb::parallel_for(NumEls, [&](int i0, int i1, int) {
                for (int i = i0; i < i1; i++)
                {
                    CMElement &el = meshele[i];
                    for (int sd = 0; sd < 3; sd++)
                    {
                        const int a = el.p[sd], b = el.p[(sd + 1) % 3];
                        if (!onTagged[a] || !onTagged[b]) continue;
                        auto f = tagged.find((long long)std::min(a, b) * NumNodes + std::max(a, b));
                        if (f != tagged.end()) el.e[sd] = f->second;
                    }
                }
            });
        }
    }
    tick("edges");
    if (deleteFiles)
        for (const char *ext : {".ele", ".node", ".pbc", ".poly"}) remove((PathName + ext).c_str());
    return NOERROR;
}

// ---------------------------------------------------------------------------------------
// .ans writers
// ---------------------------------------------------------------------------------------
namespace
{
bool echo_fem(const std::string &path, FILE *fp)
{
    FILE *fz = fopen((path + ".fem").c_str(), "rt");
    if (!fz) return false;
    char c[1024];
    while (fgets(c, 1024, fz) != NULL) fputs(c, fp);
    fclose(fz);
    return true;
}
} // namespace

static void write_tail(FSolver &S, FILE *fp, bool harmonic)
{
    // periodic pairs and air-gap tables (static2d.cpp:1150-1191 ; harmonic2d.cpp:994-1030 uses
    // a different separator in the pair rows)
    fprintf(fp, "%i\n", S.NumPBCs);
    for (int k = 0; k < S.NumPBCs; k++)
        fprintf(fp, harmonic ? "%i  %i %i\n" : "%i\t%i\t%i\n", S.pbclist[k].x, S.pbclist[k].y, S.pbclist[k].t);
    fprintf(fp, "%i\n", S.NumAirGapElems);
    for (int i = 0; i < S.NumAirGapElems; i++)
    {
        const CAirGapElement &g = S.agelist[i];
        fprintf(fp, "%s", g.BdryName.c_str());
        fprintf(fp, "%i %.17g %.17g %.17g %.17g %.17g %.17g %.17g %i %.17g %.17g\n", g.BdryFormat, g.InnerAngle,
                g.OuterAngle, g.ri, g.ro, g.totalArcLength, g.agc.re, g.agc.im, g.totalArcElements, g.InnerShift,
                g.OuterShift);
        for (int k = 0; k <= g.totalArcElements; k++)
        {
            const CQuadPoint &q = g.quadNode[k];
            fprintf(fp, "%i %.17g %i %.17g %i %.17g %i %.17g\n", q.n0, q.w0, q.n1, q.w1, q.n2, q.w2, q.n3, q.w3);
        }
    }
}

int FSolver::WriteStatic2D(CBigLinProb &L)
{
    char msg[1024];
    std::string out = PathName + ".ans";
    FILE *fp = fopen(out.c_str(), "wt");
    if (!fp)
    {
        snprintf(msg, sizeof msg, "Couldn't write to %s.ans\n", PathName.c_str());
        WarnMessage(msg);
        return false;
    }
    if (!echo_fem(PathName, fp))
    {
        fclose(fp);
        snprintf(msg, sizeof msg, "Couldn't open %s.fem\n", PathName.c_str());
        WarnMessage(msg);
        return false;
    }
    static const double unitconv[] = {2.54, 0.1, 1., 100., 0.00254, 1.e-04};
    const double cf = unitconv[LengthUnits];
    fprintf(fp, "[Solution]\n");
    fprintf(fp, "%i\n", NumNodes);
    for (int i = 0; i < NumNodes; i++)
    {
        fprintf(fp, "%.17g\t%.17g\t%.17g\t%i", meshnode[i].x / cf, meshnode[i].y / cf, L.b[i], meshnode[i].BoundaryMarker);
        if (!Aprev.empty()) fprintf(fp, "%.17g\n", Aprev[i]); else fprintf(fp, "\n");
    }
    fprintf(fp, "%i\n", NumEls);
    for (int i = 0; i < NumEls; i++)
        fprintf(fp, "%i\t%i\t%i\t%i\n", meshele[i].p[0], meshele[i].p[1], meshele[i].p[2], meshele[i].lbl);
    // circuit results per block label (static2d.cpp:1122-1148)
    fprintf(fp, "%i\n", NumBlockLabels);
    for (int k = 0; k < NumBlockLabels; k++)
    {
        int i = labellist[k].InCircuit;
        if (i < 0) fprintf(fp, "1\t0\n");
        else
        {
            if (circproplist[i].Case == 0) fprintf(fp, "0\t%.17g\n", circproplist[i].dV.Re());
            if (circproplist[i].Case == 1) fprintf(fp, "1\t%.17g\n", circproplist[i].J.Re());
        }
    }
    write_tail(*this, fp, false);
    fclose(fp);
    return true;
}

int FSolver::WriteHarmonic2D(CBigComplexLinProb &L)
{
    char msg[1024];
    std::string out = PathName + ".ans";
    FILE *fp = fopen(out.c_str(), "wt");
    if (!fp)
    {
        snprintf(msg, sizeof msg, "Couldn't write to %s.ans\n", PathName.c_str());
        WarnMessage(msg);
        return false;
    }
    if (!echo_fem(PathName, fp))
    {
        fclose(fp);
        snprintf(msg, sizeof msg, "Couldn't open %s.fem\n", PathName.c_str());
        WarnMessage(msg);
        return false;
    }
    static const double unitconv[] = {2.54, 0.1, 1., 100., 0.00254, 1.e-04};
    const double cf = unitconv[LengthUnits];
    fprintf(fp, "[Solution]\n");
    fprintf(fp, "%i\n", NumNodes);
    for (int i = 0; i < NumNodes; i++)
    {
        fprintf(fp, "%.17g\t%.17g\t%.17g\t%.17g\t%i", meshnode[i].x / cf, meshnode[i].y / cf, L.b[i].re, L.b[i].im,
                meshnode[i].BoundaryMarker);
        if (!Aprev.empty()) fprintf(fp, "\t%.17g\n", Aprev[i]); else fprintf(fp, "\n");
    }
    fprintf(fp, "%i\n", NumEls);
    for (int i = 0; i < NumEls; i++)
    {
        fprintf(fp, "%i\t%i\t%i\t%i\t%i\t%i\t%i", meshele[i].p[0], meshele[i].p[1], meshele[i].p[2], meshele[i].lbl,
                meshele[i].e[0], meshele[i].e[1], meshele[i].e[2]);
        if (!Aprev.empty()) fprintf(fp, "\t%.17g\n", meshele[i].Jprev); else fprintf(fp, "\n");
    }
    fprintf(fp, "%i\n", NumBlockLabels);
    for (int k = 0; k < NumBlockLabels; k++)
    {
        int i = labellist[k].InCircuit;
        if (i < 0) fprintf(fp, "1\t0\t0\n");
        else
        {
            if (circproplist[i].Case == 0) fprintf(fp, "0\t%.17g\t%.17g\n", circproplist[i].dV.Re(), circproplist[i].dV.Im());
            if (circproplist[i].Case == 1) fprintf(fp, "1\t%.17g\t%.17g\n", circproplist[i].J.Re(), circproplist[i].J.Im());
            if (circproplist[i].Case == 2) fprintf(fp, "0\t%.17g\t%.17g\n", L.b[NumNodes + i].Re(), L.b[NumNodes + i].Im());
        }
    }
    write_tail(*this, fp, true);
    fclose(fp);
    return true;
}
