// Host-side Model (see model.h).  Citations: /root/reference.
#include "model.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <regex>

#include "host_util.h"

namespace gopf {

using namespace parser;

// Go cmplx.Pow(x, complex(p, 0)) on the host (math/cmplx/pow.go)
std::complex<double> go_cpow_host(std::complex<double> x, double p) {
    if (x.real() == 0.0 && x.imag() == 0.0) {
        if (p == 0.0) return {1.0, 0.0};
        if (p < 0.0) return {INFINITY, 0.0};
        return {0.0, 0.0};
    }
    const double modulus = std::hypot(x.real(), x.imag());
    const double r = std::pow(modulus, p);
    const double theta = p * std::atan2(x.imag(), x.real());
    return {r * std::cos(theta), r * std::sin(theta)};
}

Model::Model() : N(0) {}

void Model::panic_on_prefix_in_name(const std::string& name) const {
    if (!get_known_prefixes(name).empty())
        throw Error("The words [- LAP^4 LAP^2 LAP *] are reserved. Do not include them in your variable names");
}

// ---- pf.Model API -------------------------------------------------------------------
void Model::add_field(const std::string& name, size_t n, double* host) {
    panic_on_prefix_in_name(name);  // NewField, model.go:44-45
    if (n == 0) throw Error("model: field '" + name + "' has no nodes");
    if (fields.empty()) N = n;
    if (fields.size() >= GOPF_MAX_FIELDS) throw Error(strf("model: at most %d fields are supported", GOPF_MAX_FIELDS));
    if (!derived.empty())
        throw Error("model: add every field before equations / functions create derived fields "
                    "(spectrum numbering is fields first)");
    fields.push_back({name, host, n});
    initialised = false;
}

void Model::add_scalar(const std::string& name, double re, double im) {
    if (scalars.find(name) == scalars.end()) scalar_order.push_back(name);
    scalars[name] = {re, im};
    initialised = false;
}

void Model::add_equation(const std::string& eq_in) {
    const std::string eq = strip_spaces(eq_in);  // model.go:158
    if (split(eq, "=").size() != 2) throw Error("build: equality sign can only occur once");
    if (fields.empty()) throw Error("Model: No fields added");
    field_name_from_leibniz(split(eq, "=")[0]);
    equations.push_back(eq);
    update_derived_fields(eq);
    sources.emplace_back();  // model.go:161
    initialised = false;
}

// model.go:151-154.  The reference indexes AllSources[eqNo] directly (a Go panic when the equation
// does not exist yet) and Source.Eval takes Dot(freq, Pos) over the frequency components
// (sourceTerm.go:28, pfutil/sliceOperations.go:52-58), so Pos needs at least `rank` entries; that is
// checked when the solver knows the rank.
void Model::add_source(int eq_no, const double* pos, int npos, SourceFn f, void* user) {
    if (eq_no < 0 || eq_no >= (int)sources.size())
        throw Error(strf("AddSource: index out of range [%d] with length %d", eq_no, (int)sources.size()));
    if (attached_solvers > 0)
        throw Error("AddSource: a solver has already been compiled from this model; add the sources before NewSolver "
                    "(the reference reads AllSources live, the device program is compiled once)");
    if (!pos || npos < 1 || npos > 3) throw Error("AddSource: Pos must hold 1..3 coordinates");
    if (!f) throw Error("AddSource: the time function is NULL");
    if ((int)sources[eq_no].size() >= GOPF_MAX_SOURCES)
        throw Error(strf("AddSource: at most %d sources per equation", GOPF_MAX_SOURCES));
    SourceSpec s;
    s.pos[0] = s.pos[1] = s.pos[2] = 0.0;
    for (int k = 0; k < npos; ++k) s.pos[k] = pos[k];
    s.npos = npos;
    s.fn = f;
    s.user = user;
    sources[eq_no].push_back(s);
    initialised = false;
}

// model.go:165-195
void Model::update_derived_fields(const std::string& eq) {
    const std::string rhs = split(eq, "=")[1];
    const std::string field = field_name_from_leibniz(split(eq, "=")[0]);
    std::vector<std::string> field_names;
    for (const HostField& f : fields) field_names.push_back(f.name);
    for (const SubStringDelimiter& sd : split_on_many(rhs, {"+", "-"})) {
        if (is_user_term(sd.SubString)) continue;
        const std::string nf = sort_factors(get_non_linear_field_expressions(sd.SubString, field, field_names));
        if (!nf.empty() && !is_field_name(nf)) register_derived_monomial(nf);
    }
}

void Model::register_derived_monomial(const std::string& desc) {
    if (is_field_name(desc)) return;  // registerDerivedFields skips known names (model.go:339-346)
    DerivedSpec d;
    d.name = desc;
    d.origin = DerivedOrigin::Monomial;
    d.dev = compile_monomial(desc);
    d.source = desc;
    d.used = false;
    derived.push_back(d);
    initialised = false;
}

void Model::register_function(const std::string& name, const std::string& expr) {
    panic_on_prefix_in_name(name);  // model.go:401
    if (fields.empty()) throw Error("Model: No fields added");
    DerivedSpec d;
    d.name = name;
    d.origin = DerivedOrigin::Function;
    d.dev = compile_expression(expr);
    d.source = expr;
    d.used = false;
    derived.push_back(d);  // RegisterDerivedField appends unconditionally (model.go:415-418)
    initialised = false;
}

void Model::register_white_noise(const std::string& name, double strength, unsigned long long seed) {
    panic_on_prefix_in_name(name);
    if (fields.empty()) throw Error("Model: No fields added");
    DerivedSpec d;
    d.name = name;
    d.origin = DerivedOrigin::WhiteNoise;
    std::memset(&d.dev, 0, sizeof(d.dev));
    d.dev.kind = DK_WHITE_NOISE;
    d.dev.noise_std = std::sqrt(2.0 * strength);  // noise.go:21
    // The reference draws every WhiteNoise instance from the shared math/rand stream, so two noise fields are
    // independent.  The Philox key is (seed, step, node): mix in the instance number (0 for the first, so a
    // model with one noise field keeps its stream) or two fields registered with the same seed -- the Python
    // mirror's default -- would be identical at every node and step.
    unsigned long long instance = 0;
    for (const DerivedSpec& o : derived)
        if (o.origin == DerivedOrigin::WhiteNoise) instance++;
    d.dev.seed = seed ^ (0x9E3779B97F4A7C15ULL * instance);
    d.source = "white_noise";
    d.used = false;
    derived.push_back(d);
    initialised = false;
}

void Model::register_table_field(const std::string& name, const double* values, long long n_steps) {
    panic_on_prefix_in_name(name);
    if (fields.empty()) throw Error("Model: No fields added");
    if (!values || n_steps < 1) throw Error("table field: needs at least one step of values");
    DerivedSpec d;
    d.name = name;
    d.origin = DerivedOrigin::Table;
    std::memset(&d.dev, 0, sizeof(d.dev));
    d.dev.kind = DK_TABLE;
    d.dev.table_steps = n_steps;
    d.table.assign(values, values + (size_t)n_steps * N);
    d.source = "table";
    d.used = false;
    derived.push_back(d);
    initialised = false;
}

void Model::register_user_term(const UserTerm& t) {
    panic_on_prefix_in_name(t.name);  // model.go:320, 372
    user_terms[t.name] = t;
    initialised = false;
}

// ---- queries ------------------------------------------------------------------------
bool Model::is_field_name(const std::string& n) const {
    for (const HostField& f : fields)
        if (f.name == n) return true;
    for (const DerivedSpec& d : derived)
        if (d.name == n) return true;
    return false;
}

bool Model::is_brick_name(const std::string& n) const { return is_field_name(n) || scalars.count(n) > 0; }

bool Model::is_user_term(const std::string& n) const { return user_terms.count(n) > 0; }

std::vector<std::string> Model::all_field_names() const {
    std::vector<std::string> out;
    for (const HostField& f : fields) out.push_back(f.name);
    for (const DerivedSpec& d : derived) out.push_back(d.name);
    return out;
}

int Model::field_index(const std::string& n) const {
    for (size_t i = 0; i < fields.size(); ++i)
        if (fields[i].name == n) return (int)i;
    return -1;
}

int Model::spectrum_index(const std::string& n) const {
    const int fi = field_index(n);
    if (fi >= 0) return fi;
    // the LAST registration wins, like m.Bricks[name] = &d (model.go:191, 417)
    for (int i = (int)derived.size() - 1; i >= 0; --i)
        if (derived[i].name == n) return (int)fields.size() + i;
    return -1;
}

int Model::eq_number(const std::string& field) const {
    static const std::regex rx("d(.*?)/dt");
    for (size_t i = 0; i < equations.size(); ++i) {
        std::smatch m;
        if (std::regex_search(equations[i], m, rx) && m.str(1) == field) return (int)i;
    }
    throw Error("EqNumber: Could not find an equation for field " + field);
}

void Model::mark_used(int spectrum) {
    const int d = spectrum - (int)fields.size();
    if (d >= 0 && d < (int)derived.size()) derived[d].used = true;
}

// ---- derived-field compilers ---------------------------------------------------------
// util.go:38-65 DerivedFieldCalcFromDesc
DevDerived Model::compile_monomial(const std::string& desc) const {
    DevDerived D;
    std::memset(&D, 0, sizeof(D));
    D.kind = DK_MONOMIAL;
    const std::vector<std::string> res = go_find_all("[^\\*]*", desc);
    if ((int)res.size() > GOPF_MAX_FACTORS)
        throw Error(strf("derived field '%s': more than %d factors", desc.c_str(), GOPF_MAX_FACTORS));
    for (const std::string& r : res) {
        const std::string name = go_find_string("^[^\\^]*", r);
        const int fi = field_index(name);
        if (fi < 0) throw Error("derived field '" + desc + "': '" + name + "' is not a field");  // nil map entry in Go
        const double p = get_power(r);
        D.field[D.n_factors] = fi;
        D.power[D.n_factors] = p;
        D.ipower[D.n_factors] = (p >= 0.0 && p <= 64.0 && p == std::floor(p)) ? (int)p : -1;
        D.n_factors++;
    }
    return D;
}

namespace {
// infix -> RPN for RegisterFunction bodies (real-valued, acts on real parts like the
// reference's closures, e.g. examples/strain_single_precipitate/main.go:45-64)
struct ExprCompiler {
    const Model& m;
    const std::string& s;
    size_t p = 0;
    DevDerived& D;
    ExprCompiler(const Model& m_, const std::string& s_, DevDerived& D_) : m(m_), s(s_), D(D_) {}

    void emit(RpnOp op, double arg = 0.0) {
        if (D.n_ops >= GOPF_MAX_RPN) throw Error("function expression too long (GOPF_MAX_RPN)");
        D.op[D.n_ops] = (unsigned char)op;
        D.arg[D.n_ops] = arg;
        D.n_ops++;
    }
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
    bool eat(char c) {
        ws();
        if (p < s.size() && s[p] == c) { ++p; return true; }
        return false;
    }
    [[noreturn]] void fail(const std::string& why) { throw Error("function expression '" + s + "': " + why + strf(" (at %zu)", p)); }

    void expr() {
        term();
        for (;;) {
            if (eat('+')) { term(); emit(OP_ADD); }
            else if (eat('-')) { term(); emit(OP_SUB); }
            else break;
        }
    }
    void term() {
        unary();
        for (;;) {
            if (eat('*')) { unary(); emit(OP_MUL); }
            else if (eat('/')) { unary(); emit(OP_DIV); }
            else break;
        }
    }
    void unary() {
        if (eat('-')) { unary(); emit(OP_NEG); return; }
        if (eat('+')) { unary(); return; }
        power();
    }
    void power() {
        primary();
        ws();
        if (eat('^')) {
            ws();
            // integer literal exponent -> repeated multiplication
            size_t q = p;
            while (q < s.size() && std::isdigit((unsigned char)s[q])) ++q;
            if (q > p && (q == s.size() || (s[q] != '.' && s[q] != 'e' && s[q] != 'E'))) {
                const int n = std::atoi(s.substr(p, q - p).c_str());
                p = q;
                emit(OP_POWI, (double)n);
            } else {
                unary();
                emit(OP_POW);
            }
        }
    }
    void primary() {
        ws();
        if (p >= s.size()) fail("unexpected end");
        if (eat('(')) {
            expr();
            if (!eat(')')) fail("missing ')'");
            return;
        }
        if (std::isdigit((unsigned char)s[p]) || s[p] == '.') {
            char* endp = nullptr;
            const double v = std::strtod(s.c_str() + p, &endp);
            if (endp == s.c_str() + p) fail("bad number");
            p = (size_t)(endp - s.c_str());
            emit(OP_CONST, v);
            return;
        }
        if (std::isalpha((unsigned char)s[p]) || s[p] == '_') {
            size_t q = p;
            while (q < s.size() && (std::isalnum((unsigned char)s[q]) || s[q] == '_')) ++q;
            const std::string id = s.substr(p, q - p);
            p = q;
            ws();
            if (p < s.size() && s[p] == '(') {
                ++p;
                if (id == "re" || id == "im") {
                    ws();
                    size_t q2 = p;
                    while (q2 < s.size() && (std::isalnum((unsigned char)s[q2]) || s[q2] == '_')) ++q2;
                    const int fi = m.field_index(s.substr(p, q2 - p));
                    if (fi < 0) fail("re()/im() need a field name");
                    p = q2;
                    if (!eat(')')) fail("missing ')'");
                    emit(id == "re" ? OP_FIELD_RE : OP_FIELD_IM, (double)fi);
                    return;
                }
                expr();
                if (!eat(')')) fail("missing ')'");
                static const std::map<std::string, RpnOp> fn = {
                    {"H", OP_H}, {"dH", OP_DH}, {"Landau", OP_LANDAU}, {"dLandau", OP_DLANDAU},
                    {"exp", OP_EXP}, {"log", OP_LOG}, {"sin", OP_SIN}, {"cos", OP_COS},
                    {"tanh", OP_TANH}, {"sqrt", OP_SQRT}, {"abs", OP_ABS}, {"negpart", OP_NEGPART}};
                auto it = fn.find(id);
                if (it == fn.end()) fail("unknown function '" + id + "'");
                emit(it->second);
                return;
            }
            const int fi = m.field_index(id);
            if (fi >= 0) { emit(OP_FIELD_RE, (double)fi); return; }
            auto sc = m.scalars.find(id);
            if (sc != m.scalars.end()) { emit(OP_CONST, sc->second.real()); return; }
            if (id == "pi") { emit(OP_CONST, 3.14159265358979323846); return; }
            fail("unknown name '" + id + "' (fields and scalars must be added first)");
        }
        fail(std::string("unexpected character '") + s[p] + "'");
    }
};
}  // namespace

// A registered function that is a real polynomial (degree <= GOPF_MAX_POLY) of re(field 0) alone -- the PFC
// ideal-mixture nonlinearity 3 a v^2 + 4 b v^3 (pf/pairCorrelationTerm.go:144-156), Landau and interpolation
// polynomials -- gets its coefficients recorded next to the program, so that the fused real-space kernels can
// evaluate it by Horner on register-resident cells (step_program.h derived_poly).  The program itself stays:
// every other path keeps interpreting (or compiling) it.
static void detect_polynomial(DevDerived* D) {
    typedef std::vector<double> Poly;
    D->poly_deg = -1;
    for (int i = 0; i <= GOPF_MAX_POLY; ++i) D->poly[i] = 0.0;
    auto mul = [](const Poly& a, const Poly& b) {
        Poly r(a.size() + b.size() - 1, 0.0);
        for (size_t i = 0; i < a.size(); ++i)
            for (size_t j = 0; j < b.size(); ++j) r[i + j] += a[i] * b[j];
        return r;
    };
    auto lin = [](double ca, const Poly& a, double cb, const Poly& b) {
        Poly r(std::max(a.size(), b.size()), 0.0);
        for (size_t i = 0; i < a.size(); ++i) r[i] += ca * a[i];
        for (size_t i = 0; i < b.size(); ++i) r[i] += cb * b[i];
        return r;
    };
    std::vector<Poly> st;
    for (int i = 0; i < D->n_ops; ++i) {
        const double arg = D->arg[i];
        switch (D->op[i]) {
            case OP_CONST: st.push_back(Poly{arg}); break;
            case OP_FIELD_RE:
                if ((int)arg != 0) return;
                st.push_back(Poly{0.0, 1.0});
                break;
            case OP_ADD: case OP_SUB: case OP_MUL: {
                if (st.size() < 2) return;
                const Poly b = st.back();
                st.pop_back();
                const Poly a = st.back();
                st.pop_back();
                st.push_back(D->op[i] == OP_MUL ? mul(a, b) : lin(1.0, a, D->op[i] == OP_ADD ? 1.0 : -1.0, b));
                break;
            }
            case OP_NEG:
                if (st.empty()) return;
                for (double& c : st.back()) c = -c;
                break;
            case OP_POWI: {
                if (st.empty() || arg < 0 || arg > GOPF_MAX_POLY || arg != std::floor(arg)) return;
                Poly r{1.0};
                for (int k = 0; k < (int)arg; ++k) r = mul(r, st.back());
                st.back() = r;
                break;
            }
            case OP_H: case OP_DH: case OP_LANDAU: case OP_DLANDAU: {
                if (st.empty()) return;
                const Poly x = st.back(), x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2);
                if (D->op[i] == OP_H) st.back() = lin(3.0, x2, -2.0, x3);
                else if (D->op[i] == OP_DH) st.back() = lin(6.0, x, -6.0, x2);
                else if (D->op[i] == OP_LANDAU) st.back() = lin(1.0, lin(1.0, x2, -2.0, x3), 1.0, x4);
                else st.back() = lin(1.0, lin(2.0, x, -6.0, x2), 4.0, x3);
                break;
            }
            default: return;  // division, transcendental functions, imaginary parts
        }
        if (!st.empty() && st.back().size() > 4 * (GOPF_MAX_POLY + 1)) return;
    }
    if (st.size() != 1) return;
    Poly r = st.back();
    while (r.size() > 1 && r.back() == 0.0) r.pop_back();
    if ((int)r.size() > GOPF_MAX_POLY + 1) return;
    for (size_t i = 0; i < r.size(); ++i) D->poly[i] = r[i];
    D->poly_deg = (int)r.size() - 1;
}

DevDerived Model::compile_expression(const std::string& expr) const {
    DevDerived D;
    std::memset(&D, 0, sizeof(D));
    D.kind = DK_RPN;
    ExprCompiler c(*this, expr, D);
    c.expr();
    c.ws();
    if (c.p != expr.size()) c.fail("trailing characters");
    // stack-depth check
    int sp = 0, maxsp = 0;
    for (int i = 0; i < D.n_ops; ++i) {
        switch (D.op[i]) {
            case OP_CONST: case OP_FIELD_RE: case OP_FIELD_IM: sp++; break;
            case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_POW: sp--; break;
            default: break;
        }
        maxsp = std::max(maxsp, sp);
    }
    if (sp != 1 || maxsp > GOPF_RPN_STACK) throw Error("function expression '" + expr + "' is too deeply nested");
    detect_polynomial(&D);
    return D;
}

// ---- Build ---------------------------------------------------------------------------
// rhsBuilder.go:109-122 ValidName
static bool valid_name(const Model& m, const std::string& name) {
    std::string s = strip_spaces(name);
    if (s.empty() || s == "LAP") return true;
    if (s.size() >= 3 && s.substr(0, 3) == "LAP") s = s.substr(3);
    return m.is_brick_name(s);
}

// rhsBuilder.go:125-190 ConcreteTerm -> coef * [brick] * L^lap
DevTerm Model::concrete_term(const SubStringDelimiter& td) const {
    const std::string& term = td.SubString;
    std::complex<double> coef(td.PreceedingDelimiter == "-" ? -1.0 : 1.0, 0.0);
    for (const std::string& r : go_find_all("[^\\*]*", term)) {
        const std::string name = go_find_string("^[^\\^]*", r);
        if (!is_field_name(name) && is_brick_name(name)) {
            coef *= go_cpow_host(scalars.at(name), get_power(r));
        } else if (!valid_name(*this, name)) {
            throw Error("rhsBuilder: Name " + name + " is not defined!");
        }
    }
    const std::string field_name = get_field_name(sort_factors(term), all_field_names());
    DevTerm t;
    std::memset(&t, 0, sizeof(t));
    t.cre = coef.real();
    t.cim = coef.imag();
    t.kind = TK_MONOMIAL;
    t.brick = field_name.empty() ? -1 : spectrum_index(field_name);
    t.lap = 0;
    if (contains(term, "LAP")) t.lap = (int)get_power(go_find_string("LAP*[^a-zA-Z]*", term));
    return t;
}

// rhsBuilder.go:199-242 constructFunc
void Model::apply_prefixes(DevTerm* t, const std::vector<std::string>& prefixes) const {
    for (const std::string& p : prefixes) {
        if (p == "-") { t->cre = -t->cre; t->cim = -t->cim; }
        else if (p == "LAP^4") t->lap += 4;
        else if (p == "LAP^2") t->lap += 2;
        else if (p == "LAP") t->lap += 1;
        // " ", "+" : nothing; anything else: the reference logs 'Unrecognized prefix' and keeps the term
    }
}

CompiledEquation Model::build(const std::string& eq) {
    const std::vector<std::string> sides = split(eq, "=");
    if (sides.size() != 2) throw Error("build: equality sign can only occur once");
    CompiledEquation ce;
    ce.text = eq;
    ce.field = field_name_from_leibniz(sides[0]);
    for (SubStringDelimiter t : split_on_many(sides[1], {"+", "-"})) {
        const std::string name = remove_known_prefixes(t.SubString);
        std::vector<std::string> prefixes = get_known_prefixes(t.SubString);
        prefixes.push_back(t.PreceedingDelimiter);
        auto ut = user_terms.find(name);
        if (ut != user_terms.end()) {
            UserTerm& u = ut->second;
            DevTerm d;
            std::memset(&d, 0, sizeof(d));
            d.cre = 1.0;
            d.brick = -1;
            d.param = u.slot;
            switch (u.kind) {
                case UserTermKind::SpectralViscosity:
                    d.kind = TK_SPECTRAL_VISC;
                    apply_prefixes(&d, prefixes);
                    ce.den.push_back(d);
                    break;
                case UserTermKind::TensorialHessian:
                    d.kind = TK_TENSOR_HESSIAN;
                    apply_prefixes(&d, prefixes);
                    ce.den.push_back(d);
                    break;
                case UserTermKind::PairCorrelation:
                    d.kind = TK_PAIR_CORR;
                    d.lap = u.laplacian ? 1 : 0;
                    apply_prefixes(&d, prefixes);
                    ce.den.push_back(d);
                    break;
                case UserTermKind::ExplicitPairCorrelation:
                    d.kind = TK_PAIR_CORR;
                    d.lap = u.laplacian ? 1 : 0;
                    d.brick = spectrum_index(u.field);
                    if (d.brick < 0) throw Error("ExplicitPairCorrelationTerm: unknown field " + u.field);
                    mark_used(d.brick);
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
                case UserTermKind::IdealMixture: {
                    // pairCorrelationTerm.go:126-178: linear part Prefactor [*LAP], non-linear part the
                    // derived field ideal_mixture_<field>_nonlin [*LAP]
                    DevTerm lin = d;
                    lin.kind = TK_MONOMIAL;
                    lin.cre = u.prefactor;
                    lin.lap = u.laplacian ? 1 : 0;
                    apply_prefixes(&lin, prefixes);
                    ce.den.push_back(lin);
                    DevTerm nl = d;
                    nl.kind = TK_MONOMIAL;
                    const std::string dn = "ideal_mixture_" + u.field + "_nonlin";
                    nl.brick = spectrum_index(dn);
                    if (nl.brick < 0)
                        throw Error("Missing derived field " + dn +
                                    ".\nMake sure that the field returned by IdealMixtureTerm.DerivedField is registered");
                    mark_used(nl.brick);
                    nl.lap = u.laplacian ? 1 : 0;
                    apply_prefixes(&nl, prefixes);
                    ce.rhs.push_back(nl);
                    break;
                }
                case UserTermKind::ConservativeNoise:
                    d.kind = TK_CONS_NOISE;
                    for (const std::string& cn : u.current_names) {
                        const int si = spectrum_index(cn);
                        if (si < 0)
                            throw Error("ConservativeCurrent: Current fields are not register. Make sure that you have "
                                        "registered the fields returned by RequiredDerivedFields.");
                        mark_used(si);
                    }
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
                case UserTermKind::VolumeConservingLP:
                    d.kind = TK_VOLUME_LP;
                    d.brick = spectrum_index(u.indicator);
                    if (d.brick < 0) throw Error("VolumeConservingLP: Indicator is not a derived field");
                    mark_used(d.brick);
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
                case UserTermKind::SquaredGradient:
                    d.kind = TK_MONOMIAL;
                    d.brick = u.work_spectrum;
                    d.cre = u.prefactor;
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
                case UserTermKind::ChargeTransport:
                    // the solver fills the work spectrum with the whole term (chargeTransport.go:94-119)
                    if (field_index(u.field) < 0) throw Error("ChargeTransport: unknown field " + u.field);
                    if (u.conductivity.size() != (size_t)u.n_voigt * N)
                        throw Error(strf("ChargeTransport: conductivity table holds %zu values, expected %d x %zu",
                                         u.conductivity.size(), u.n_voigt, N));
                    d.kind = TK_MONOMIAL;
                    d.brick = u.work_spectrum;
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
                case UserTermKind::HomogeneousModulusLinElast:
                    // the solver fills the work spectrum with the whole term (homoLinElast.go:47-99)
                    if (field_index(u.field) < 0) throw Error("HomogeneousModulusLinElast: unknown field " + u.field);
                    d.kind = TK_MONOMIAL;
                    d.brick = u.work_spectrum;
                    apply_prefixes(&d, prefixes);
                    ce.rhs.push_back(d);
                    break;
            }
        } else if (is_bilinear(t.SubString, ce.field, all_field_names())) {
            t.SubString = replace_all(t.SubString, ce.field, "");
            ce.den.push_back(concrete_term(t));
        } else {
            DevTerm d = concrete_term(t);
            const int di = d.brick - (int)fields.size();
            if (kspace_noise && di >= 0 && di < (int)derived.size() && derived[di].dev.kind == DK_WHITE_NOISE) {
                // + c * L^n * NOISE: the noise spectrum is drawn at the k-point, the field is never transformed
                int slot = -1;
                for (const auto& ks : knoise_slots)
                    if (ks.second == di) slot = ks.first;
                if (slot < 0) {
                    slot = knoise_base_ + (int)knoise_slots.size();
                    if (slot >= GOPF_MAX_SPECIAL)
                        throw Error(strf("model: at most %d TensorialHessian terms and k-space noise fields together", GOPF_MAX_SPECIAL));
                    knoise_slots.push_back({slot, di});
                }
                d.kind = TK_WHITE_NOISE_K;
                d.brick = -1;
                d.param = slot;
            }
            if (d.brick >= 0) mark_used(d.brick);
            ce.rhs.push_back(d);
        }
    }
    if ((int)ce.rhs.size() > GOPF_MAX_TERMS || (int)ce.den.size() > GOPF_MAX_TERMS)
        throw Error(strf("equation '%s': more than %d terms on one side", eq.c_str(), GOPF_MAX_TERMS));
    return ce;
}

// model.go:244-260
void Model::init() {
    if (fields.empty()) throw Error("Model: No fields added");
    if (equations.size() > fields.size()) throw Error("model: more equations than fields");
    for (DerivedSpec& d : derived) d.used = false;
    // assign parameter slots / work spectra to the registered user terms
    int n_sv = 0, n_pc = 0, n_cn = 0, n_lp = 0, n_el = 0, n_th = 0, n_ct = 0;
    n_work_spectra = 0;
    for (auto& kv : user_terms) {
        UserTerm& u = kv.second;
        switch (u.kind) {
            case UserTermKind::SpectralViscosity: u.slot = n_sv++; break;
            case UserTermKind::TensorialHessian: u.slot = n_th++; break;
            case UserTermKind::PairCorrelation:
            case UserTermKind::ExplicitPairCorrelation: u.slot = n_pc++; break;
            case UserTermKind::ConservativeNoise: u.slot = n_cn++; break;
            case UserTermKind::VolumeConservingLP: u.slot = n_lp++; break;
            case UserTermKind::HomogeneousModulusLinElast:
                u.slot = n_el++;
                u.work_spectrum = (int)(fields.size() + derived.size()) + n_work_spectra++;
                break;
            case UserTermKind::ChargeTransport:
                u.slot = n_ct++;
                u.work_spectrum = (int)(fields.size() + derived.size()) + n_work_spectra++;
                break;
            case UserTermKind::SquaredGradient:
                u.work_spectrum = (int)(fields.size() + derived.size()) + n_work_spectra++;
                break;
            default: break;
        }
    }
    // one work spectrum per equation that has point sources (model.go:291-294)
    source_spectrum.assign(equations.size(), -1);
    for (size_t e = 0; e < equations.size() && e < sources.size(); ++e)
        if (!sources[e].empty()) source_spectrum[e] = (int)(fields.size() + derived.size()) + n_work_spectra++;
    if (n_sv > GOPF_MAX_SPECIAL || n_pc > GOPF_MAX_SPECIAL || n_cn > GOPF_MAX_SPECIAL || n_lp > GOPF_MAX_SPECIAL ||
        n_el > GOPF_MAX_SPECIAL || n_th > GOPF_MAX_SPECIAL || n_ct > GOPF_MAX_SPECIAL)
        throw Error(strf("model: at most %d terms of each special kind", GOPF_MAX_SPECIAL));
    if (n_spectra() > GOPF_MAX_SPECTRA) throw Error(strf("model: at most %d spectra", GOPF_MAX_SPECTRA));
    compiled.clear();
    knoise_slots.clear();
    knoise_base_ = n_th;  // k-space noise borrows the TensorHessianParams slots after the real ones
    for (const std::string& eq : equations) compiled.push_back(build(eq));
    // equation i must evolve field i (Euler.Step pairs m.RHS[i] with m.Fields[i], euler.go:27-31)
    initialised = true;
}

void Model::fill_program(DevKProgram* P, double dt, int rank) const {
    if (!initialised) throw Error("model: Init() has not been called");
    std::memset(P, 0, sizeof(*P));
    P->rank = rank;
    P->n_fields = (int)fields.size();
    P->dt = dt;
    for (size_t i = 0; i < fields.size(); ++i) {
        DevEquation& q = P->eq[i];
        if (i < compiled.size()) {
            q.n_rhs = (int)compiled[i].rhs.size();
            q.n_den = (int)compiled[i].den.size();
            for (int j = 0; j < q.n_rhs; ++j) q.rhs[j] = compiled[i].rhs[j];
            for (int j = 0; j < q.n_den; ++j) q.den[j] = compiled[i].den[j];
            if (i < source_spectrum.size() && source_spectrum[i] >= 0) {  // + sum of the sources (model.go:291-294)
                if (q.n_rhs >= GOPF_MAX_TERMS)
                    throw Error(strf("equation %d: more than %d terms on one side", (int)i, GOPF_MAX_TERMS));
                DevTerm t;
                t.cre = 1.0;
                t.cim = 0.0;
                t.brick = source_spectrum[i];
                t.lap = 0;
                t.kind = TK_MONOMIAL;
                t.param = 0;
                q.rhs[q.n_rhs++] = t;
            }
        }
    }
    for (const auto& ks : knoise_slots) {  // step_program.h TensorHessianParams: amplitude, seed bits, step bits
        TensorHessianParams& h = P->th[ks.first];
        const DevDerived& d = derived[ks.second].dev;
        h.K[0] = d.noise_std * std::sqrt((double)N);
        h.K[1] = gopf_double_of(d.seed);
        h.K[2] = gopf_double_of(0ull);
    }
    for (const auto& kv : user_terms) {
        const UserTerm& u = kv.second;
        if (u.slot < 0) continue;
        switch (u.kind) {
            case UserTermKind::SpectralViscosity: P->sv[u.slot] = u.sv; break;
            case UserTermKind::TensorialHessian: P->th[u.slot] = u.hessian; break;
            case UserTermKind::PairCorrelation:
            case UserTermKind::ExplicitPairCorrelation:
                P->pc[u.slot] = u.pc;
                P->pc[u.slot].prefactor = u.prefactor;
                break;
            case UserTermKind::ConservativeNoise: {
                ConsNoiseParams c;
                std::memset(&c, 0, sizeof(c));
                c.dim = u.dim;
                for (int k = 0; k < u.dim && k < 3; ++k) c.brick[k] = spectrum_index(u.current_names[k]);
                P->cn[u.slot] = c;
                break;
            }
            default: break;
        }
    }
}

}  // namespace gopf
