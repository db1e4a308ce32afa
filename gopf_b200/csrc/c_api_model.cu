// C ABI, part 2: pf.Model / pf.Solver.  Contract: include/gopf_cuda.h.
#include <cmath>
#include <cstring>

#include "../../include/gopf_cuda.h"
#include "c_api_types.h"
#include "solver.h"

using namespace gopf;

struct gopf_solver {
    Solver* s;
    gopf_model* owner;
    std::vector<KernelTimer> prof;
};

static const char* need(const char* p, const char* what) {
    if (!p) throw Error(std::string(what) + " is NULL");
    return p;
}

static void copy_name(const std::string& s, char* buf, int len) {
    if (!buf || len <= 0) throw Error("name buffer is NULL/empty");
    std::strncpy(buf, s.c_str(), (size_t)len - 1);
    buf[len - 1] = '\0';
}

extern "C" {

int gopf_model_create(gopf_model** out) {
    GOPF_API_BEGIN
    if (!out) throw Error("gopf_model_create: out is NULL");
    *out = new gopf_model;
    GOPF_API_END
}

int gopf_model_add_field(gopf_model* m, const char* name, int64_t n_nodes, double* host) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (n_nodes <= 0) throw Error("model: Inconsistent length of data");
    m->m.add_field(need(name, "name"), (size_t)n_nodes, host);
    GOPF_API_END
}

int gopf_model_add_scalar(gopf_model* m, const char* name, double re, double im) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.add_scalar(need(name, "name"), re, im);
    GOPF_API_END
}

int gopf_model_add_equation(gopf_model* m, const char* eq) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.add_equation(need(eq, "equation"));
    GOPF_API_END
}

int gopf_model_register_function(gopf_model* m, const char* name, const char* expr) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.register_function(need(name, "name"), need(expr, "expression"));
    GOPF_API_END
}

static const DevDerived& registered_function(gopf_model* m, const char* name) {
    if (!m) throw Error("model is NULL");
    const std::string n = need(name, "name");
    for (const DerivedSpec& d : m->m.derived)
        if (d.name == n) {
            if (d.dev.kind != DK_RPN) throw Error("'" + n + "' is not a registered function");
            return d.dev;
        }
    throw Error("no derived field named '" + n + "'");
}

int gopf_model_function_source(gopf_model* m, const char* name, int kernel, char* buf, int64_t len, int64_t* needed) {
    GOPF_API_BEGIN
    const DevDerived& D = registered_function(m, name);
    const std::string src = kernel ? jit::derived_kernel_source(D, nullptr) : jit::expression_source(D, nullptr);
    if (needed) *needed = (int64_t)src.size() + 1;
    if (buf) {
        if (len < (int64_t)src.size() + 1) throw Error("gopf_model_function_source: buffer too small");
        std::memcpy(buf, src.c_str(), src.size() + 1);
    }
    GOPF_API_END
}

int gopf_model_function_pass_source(gopf_model* m, const char* name, int line_length, char* buf, int64_t len, int64_t* needed) {
    GOPF_API_BEGIN
    const DevDerived& D = registered_function(m, name);
    const std::string src = jit::derived_pass_source(D, line_length, nullptr);
    if (needed) *needed = (int64_t)src.size() + 1;
    if (buf) {
        if (len < (int64_t)src.size() + 1) throw Error("gopf_model_function_pass_source: buffer too small");
        std::memcpy(buf, src.c_str(), src.size() + 1);
    }
    GOPF_API_END
}

int gopf_model_function_pass_compile(gopf_model* m, const char* name, int line_length, int64_t* cubin_bytes, char* lowered_name,
                                     int lowered_len) {
    GOPF_API_BEGIN
    const DevDerived& D = registered_function(m, name);
    std::vector<char> cubin;
    std::string log, name_expr, lowered;
    const std::string src = jit::derived_pass_source(D, line_length, &name_expr);
    if (!jit::compile_cubin(src, &cubin, &log, &name_expr, &lowered)) throw Error("jit: " + log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    if (lowered_name) copy_name(lowered, lowered_name, lowered_len);
    GOPF_API_END
}

int gopf_model_function_compile(gopf_model* m, const char* name, int64_t* cubin_bytes) {
    GOPF_API_BEGIN
    const DevDerived& D = registered_function(m, name);
    std::vector<char> cubin;
    std::string log;
    if (!jit::compile_cubin(jit::derived_kernel_source(D, nullptr), &cubin, &log)) throw Error("jit: " + log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    GOPF_API_END
}

static std::string kupdate_source_of(gopf_model* m, int rank, const int* n, double dt, unsigned tab_mask,
                                    uint64_t filter_addr, int filter_n, uint64_t lp_addr) {
    if (!m || !n) throw Error("NULL argument");
    if (rank != 2 && rank != 3) throw Error("rank must be 2 or 3");
    m->m.init();
    DevKProgram P;
    m->m.fill_program(&P, dt, rank);
    if (filter_addr) {
        P.filter = reinterpret_cast<const double*>(uintptr_t(filter_addr));
        P.filter_n = filter_n;
    }
    // the solver keeps {multiplier, integral, first flag} per VolumeConservingLP slot (solver.cu)
    if (lp_addr)
        for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) P.lp_multiplier[i] = reinterpret_cast<const double*>(uintptr_t(lp_addr)) + 3 * i;
    FreqGeom fg;
    fg.rank = rank;
    fg.d0 = n[0];
    fg.d1 = n[1];
    fg.d2 = rank > 2 ? n[2] : 1;
    long long N = 1;
    for (int i = 0; i < rank; ++i) N *= n[i];
    return jit::kupdate_kernel_source(P, fg, N, tab_mask);
}

int gopf_model_kupdate_source(gopf_model* m, int rank, const int* n, double dt, unsigned tab_mask, uint64_t filter_addr,
                              int filter_n, uint64_t lp_addr, char* buf, int64_t len, int64_t* needed) {
    GOPF_API_BEGIN
    const std::string src = kupdate_source_of(m, rank, n, dt, tab_mask, filter_addr, filter_n, lp_addr);
    if (needed) *needed = (int64_t)src.size() + 1;
    if (buf) {
        if (len < (int64_t)src.size() + 1) throw Error("gopf_model_kupdate_source: buffer too small");
        std::memcpy(buf, src.c_str(), src.size() + 1);
    }
    GOPF_API_END
}

int gopf_model_kupdate_compile(gopf_model* m, int rank, const int* n, double dt, unsigned tab_mask, uint64_t filter_addr,
                               int filter_n, uint64_t lp_addr, int64_t* cubin_bytes) {
    GOPF_API_BEGIN
    std::vector<char> cubin;
    std::string log;
    if (!jit::compile_cubin(kupdate_source_of(m, rank, n, dt, tab_mask, filter_addr, filter_n, lp_addr), &cubin, &log))
        throw Error("jit: " + log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    GOPF_API_END
}

// raw program images: see the header
static void copy_image(const void* src, size_t size, void* buf, int64_t len, int64_t* needed) {
    if (needed) *needed = (int64_t)size;
    if (buf) {
        if (len < (int64_t)size) throw Error("image buffer too small");
        std::memcpy(buf, src, size);
    }
}

int gopf_model_program_image(gopf_model* m, int rank, double dt, void* buf, int64_t len, int64_t* needed) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (rank != 2 && rank != 3) throw Error("rank must be 2 or 3");
    m->m.init();
    DevKProgram P;
    m->m.fill_program(&P, dt, rank);
    copy_image(&P, sizeof(P), buf, len, needed);
    GOPF_API_END
}

int gopf_model_fused_program_image(gopf_model* m, int rank, double dt, void* buf, int64_t len, int64_t* needed, int* derived_index) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (rank != 2 && rank != 3) throw Error("rank must be 2 or 3");
    m->m.init();
    const int d = single_field_derived_index(m->m);
    if (d < 0) throw Error("model: not a single-field model with one nonlinearity (the fused kernels do not apply)");
    DevKProgram P;
    m->m.fill_program(&P, dt, rank);
    finalize_single_field_program(&P, (int)m->m.fields.size());
    copy_image(&P, sizeof(P), buf, len, needed);
    if (derived_index) *derived_index = d;
    GOPF_API_END
}

int gopf_model_derived_image(gopf_model* m, int index, void* buf, int64_t len, int64_t* needed, int* used) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.init();
    if (index < 0 || index >= (int)m->m.derived.size()) throw Error("derived field index out of range");
    DevDerived D = m->m.derived[index].dev;
    D.table = nullptr;  // a device address when a solver is attached; the caller supplies its own
    copy_image(&D, sizeof(D), buf, len, needed);
    if (used) *used = m->m.derived[index].used ? 1 : 0;
    GOPF_API_END
}

int gopf_model_register_white_noise(gopf_model* m, const char* name, double strength, uint64_t seed) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.register_white_noise(need(name, "name"), strength, seed);
    GOPF_API_END
}

int gopf_model_register_table_field(gopf_model* m, const char* name, const double* values, int64_t n_steps) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.register_table_field(need(name, "name"), values, n_steps);
    GOPF_API_END
}

int gopf_model_register_spectral_viscosity(gopf_model* m, const char* name, double eps, double threshold, int power) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Implicit;
    u.kind = UserTermKind::SpectralViscosity;
    u.sv.eps = eps;
    u.sv.threshold = threshold;
    u.sv.power = power;
    u.sv.pad = 0;
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_pair_correlation(gopf_model* m, const char* name, int explicit_term, const char* field,
                                         double prefactor, int laplacian, double eff_temp, int n_peaks,
                                         const double* plane_density, const double* location, const double* width,
                                         const int* num_planes) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (n_peaks < 0 || n_peaks > GOPF_MAX_PEAKS) throw Error(strf("pair correlation: at most %d peaks", GOPF_MAX_PEAKS));
    if (n_peaks > 0 && (!plane_density || !location || !width || !num_planes)) throw Error("pair correlation: NULL peak array");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = explicit_term ? UserTermClass::Explicit : UserTermClass::Implicit;
    u.kind = explicit_term ? UserTermKind::ExplicitPairCorrelation : UserTermKind::PairCorrelation;
    u.field = field ? field : "";
    u.prefactor = prefactor;
    u.laplacian = laplacian != 0;
    std::memset(&u.pc, 0, sizeof(u.pc));
    u.pc.prefactor = prefactor;
    u.pc.eff_temp = eff_temp;
    u.pc.n_peaks = n_peaks;
    for (int i = 0; i < n_peaks; ++i) {
        u.pc.plane_density[i] = plane_density[i];
        u.pc.location[i] = location[i];
        u.pc.width[i] = width[i];
        u.pc.num_planes[i] = (double)num_planes[i];
    }
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_ideal_mixture(gopf_model* m, const char* name, const char* field, double c3, double c4,
                                      double prefactor, int laplacian, int register_derived) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Mixed;
    u.kind = UserTermKind::IdealMixture;
    u.field = need(field, "field");
    u.prefactor = prefactor;
    u.laplacian = laplacian != 0;
    u.c3 = c3;
    u.c4 = c4;
    m->m.register_user_term(u);
    if (register_derived) {
        // IdealMixtureTerm.DerivedField (pairCorrelationTerm.go:144-156): 3*c3'*v*v + 4*c4'*v*v*v,
        // c3' = -C3/6, c4' = C4/12 (pfc/ideal.go:42-50)
        const std::string dn = "ideal_mixture_" + u.field + "_nonlin";
        if (!m->m.is_field_name(dn)) {
            const std::string v = "re(" + u.field + ")";
            const std::string e = strf("3.0*(%.17g)*", -c3 / 6.0) + v + "*" + v + strf("+4.0*(%.17g)*", c4 / 12.0) + v +
                                  "*" + v + "*" + v;
            m->m.register_function(dn, e);
        }
    }
    GOPF_API_END
}

static void add_cons_noise_term(gopf_model* m, const char* name, int dim, uint32_t unique_prefix) {
    if (dim < 1 || dim > 3) throw Error("ConservativeNoise: Dim must be 1..3");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Explicit;
    u.kind = UserTermKind::ConservativeNoise;
    u.dim = dim;
    for (int c = 0; c < dim; ++c) u.current_names.push_back(strf("%u_current_%d", unique_prefix, c));  // noise.go:44-46
    m->m.register_user_term(u);
}

int gopf_model_register_conservative_noise(gopf_model* m, const char* name, double strength, int dim,
                                           uint32_t unique_prefix, uint64_t seed) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    add_cons_noise_term(m, name, dim, unique_prefix);
    for (int c = 0; c < dim; ++c) {  // RequiredDerivedFields (noise.go:85-100)
        const std::string dn = strf("%u_current_%d", unique_prefix, c);
        if (!m->m.is_field_name(dn)) {
            // names start with digits; bypass the reserved-prefix check exactly like the reference (no check there)
            m->m.register_white_noise(dn, strength, seed + 0x9E3779B97F4A7C15ull * (uint64_t)(c + 1));
        }
    }
    GOPF_API_END
}

int gopf_model_register_conservative_noise_term(gopf_model* m, const char* name, int dim, uint32_t unique_prefix) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    add_cons_noise_term(m, name, dim, unique_prefix);
    GOPF_API_END
}

int gopf_model_register_volume_conserving_lp(gopf_model* m, const char* name, const char* field,
                                             const char* indicator, double dt) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Explicit;
    u.kind = UserTermKind::VolumeConservingLP;
    u.field = need(field, "field");
    u.indicator = need(indicator, "indicator");
    u.dt = dt;
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_squared_gradient(gopf_model* m, const char* name, const char* field, double factor) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Explicit;
    u.kind = UserTermKind::SquaredGradient;
    u.field = need(field, "field");
    u.prefactor = factor;
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_tensorial_hessian(gopf_model* m, const char* name, const char* field, const double* k, int n_coeff) {
    GOPF_API_BEGIN
    if (!m || !k) throw Error("NULL argument");
    if (n_coeff != 4 && n_coeff != 9) throw Error("TensorialHessian: K must hold 4 (2-D) or 9 (3-D) coefficients");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Implicit;
    u.kind = UserTermKind::TensorialHessian;
    u.field = field ? field : "";
    u.hessian.d = n_coeff == 9 ? 3 : 2;
    for (int i = 0; i < n_coeff; ++i) u.hessian.K[i] = k[i];
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_homogeneous_modulus_lin_elast(gopf_model* m, const char* name, const char* field,
                                                      const double* stiffness81, const double* misfit9) {
    GOPF_API_BEGIN
    if (!m || !stiffness81 || !misfit9) throw Error("NULL argument");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Explicit;
    u.kind = UserTermKind::HomogeneousModulusLinElast;
    u.field = need(field, "field");
    std::memcpy(u.stiffness, stiffness81, sizeof(u.stiffness));
    std::memcpy(u.misfit, misfit9, sizeof(u.misfit));
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_register_charge_transport(gopf_model* m, const char* name, const char* field, const double* conductivity,
                                         int n_voigt, int64_t n_nodes, const double* external_field, int n_ext) {
    GOPF_API_BEGIN
    if (!m || !conductivity || !external_field) throw Error("NULL argument");
    if (n_voigt != 3 && n_voigt != 6)
        throw Error("ChargeTransport: the conductivity has 3 (2-D) or 6 (3-D) Voigt components");
    const int dim = n_voigt == 3 ? 2 : 3;
    if (n_ext < dim) throw Error(strf("ChargeTransport: ExternalField needs %d components", dim));
    if (n_nodes <= 0) throw Error("ChargeTransport: empty conductivity table");
    UserTerm u;
    u.name = need(name, "name");
    u.cls = UserTermClass::Explicit;
    u.kind = UserTermKind::ChargeTransport;
    u.field = need(field, "field");
    u.n_voigt = n_voigt;
    u.conductivity.assign(conductivity, conductivity + (size_t)n_voigt * (size_t)n_nodes);
    for (int k = 0; k < dim; ++k) u.external_field[k] = external_field[k];
    m->m.register_user_term(u);
    GOPF_API_END
}

int gopf_model_add_source(gopf_model* m, int eq_no, const double* pos, int n_pos, gopf_time_fn f, void* user) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.add_source(eq_no, pos, n_pos, f, user);
    GOPF_API_END
}

int gopf_charge_transport_multipliers(int rank, const double* freq, int64_t count, double* field_mult, double* div_mult) {
    GOPF_API_BEGIN
    if (!freq) throw Error("NULL argument");
    if (rank != 2 && rank != 3) throw Error("gopf_charge_transport_multipliers: rank must be 2 or 3");
    for (int64_t i = 0; i < count; ++i)
        for (int c = 0; c < rank; ++c) {
            if (field_mult) field_mult[(int64_t)c * count + i] = ct_field_multiplier(freq + (size_t)rank * i, rank, c);
            if (div_mult) div_mult[(int64_t)c * count + i] = ct_divergence_multiplier(freq + (size_t)rank * i, c);
        }
    GOPF_API_END
}

int gopf_charge_transport_voigt_index(int i, int j, int dim, int* out) {
    GOPF_API_BEGIN
    if (!out) throw Error("NULL argument");
    if ((dim != 2 && dim != 3) || i < 0 || j < 0 || i >= dim || j >= dim) throw Error("voigtIndex: index out of range");
    *out = ct_voigt(i, j, dim);
    GOPF_API_END
}

int gopf_source_eval(int rank, const double* freq, int64_t count, const double* pos, double amp, double* out) {
    GOPF_API_BEGIN
    if (!freq || !pos || !out) throw Error("NULL argument");
    if (rank < 1 || rank > 3) throw Error("gopf_source_eval: rank must be 1..3");
    for (int64_t i = 0; i < count; ++i) source_value(freq + (size_t)rank * i, pos, rank, amp, &out[2 * i], &out[2 * i + 1]);
    GOPF_API_END
}

// ---- elasticity package helpers (host) ---------------------------------------------------
static inline int r4(int i, int j, int k, int l) { return i * 27 + j * 9 + k * 3 + l; }

// elasticity/rank4.go:113-128 (as written: only C_ijij and C_jiji carry c44)
int gopf_elasticity_cubic_material(double c11, double c12, double c44, double* out) {
    GOPF_API_BEGIN
    if (!out) throw Error("NULL argument");
    for (int i = 0; i < 81; ++i) out[i] = 0.0;
    for (int i = 0; i < 3; ++i) out[r4(i, i, i, i)] = c11;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) {
            out[r4(i, i, j, j)] = c12;
            out[r4(j, j, i, i)] = c12;
            out[r4(i, j, i, j)] = c44;
            out[r4(j, i, j, i)] = c44;
        }
    GOPF_API_END
}

// elasticity/rank4.go:78-110
int gopf_elasticity_isotropic(double bulk_mod, double poisson, double* out) {
    GOPF_API_BEGIN
    if (!out) throw Error("NULL argument");
    const double shear = 3.0 * bulk_mod * (1.0 - 2.0 * poisson) / (2.0 * (1.0 + poisson));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) {
                    double v = 0.0;
                    if (i == j && k == l) v += bulk_mod - 2.0 * shear / 3.0;
                    if (i == k && j == l) v += shear;
                    if (i == l && j == k) v += shear;
                    out[r4(i, j, k, l)] = v;
                }
    GOPF_API_END
}

// elasticity/rank4.go:39-60
int gopf_elasticity_rotate(double* c, const double* rot) {
    GOPF_API_BEGIN
    if (!c || !rot) throw Error("NULL argument");
    double res[81];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) {
                    double s = 0.0;
                    for (int mm = 0; mm < 3; ++mm)
                        for (int n = 0; n < 3; ++n)
                            for (int p = 0; p < 3; ++p)
                                for (int q = 0; q < 3; ++q)
                                    s += rot[i * 3 + mm] * rot[j * 3 + n] * rot[k * 3 + p] * rot[l * 3 + q] * c[r4(mm, n, p, q)];
                    res[r4(i, j, k, l)] = s;
                }
    std::memcpy(c, res, sizeof(res));
    GOPF_API_END
}

// elasticity/rank4.go:62-75
int gopf_elasticity_contract_last(const double* c, const double* t, double* out) {
    GOPF_API_BEGIN
    if (!c || !t || !out) throw Error("NULL argument");
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) s += c[r4(i, j, k, l)] * t[k * 3 + l];
            out[i * 3 + j] = s;
        }
    GOPF_API_END
}

// elasticity/linearElasticity.go:86-98
int gopf_elasticity_energy_density(const double* c, const double* e, double* out) {
    GOPF_API_BEGIN
    if (!c || !e || !out) throw Error("NULL argument");
    double res = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) res += c[r4(i, j, k, l)] * e[i * 3 + j] * e[k * 3 + l];
    *out = 0.5 * res;
    GOPF_API_END
}

int gopf_elasticity_multiplier(const double* c, const double* misfit, int dim, const double* freq3, int64_t count,
                               double* out) {
    GOPF_API_BEGIN
    if (!c || !misfit || !freq3 || !out) throw Error("NULL argument");
    if (dim != 2 && dim != 3) throw Error("gopf_elasticity_multiplier: dim must be 2 or 3");
    ElastParams E;
    make_elast_params(&E, c, misfit, dim);
    for (int64_t i = 0; i < count; ++i) out[i] = elastic_multiplier(E, freq3[3 * i], freq3[3 * i + 1], freq3[3 * i + 2]);
    GOPF_API_END
}

int gopf_has_kspace_noise(void) {
#ifdef GOPF_KNOISE
    return 1;
#else
    return 0;
#endif
}

int gopf_model_set_kspace_noise(gopf_model* m, int on) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (m->m.attached_solvers > 0) throw Error("model: a solver has been compiled from this model already");
    m->m.kspace_noise = on != 0;
    m->m.initialised = false;
    GOPF_API_END
}

int gopf_model_init(gopf_model* m) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    m->m.init();
    GOPF_API_END
}

int gopf_model_num_fields(gopf_model* m, int* n) {
    GOPF_API_BEGIN
    if (!m || !n) throw Error("NULL argument");
    *n = (int)m->m.fields.size();
    GOPF_API_END
}

int gopf_model_num_derived_fields(gopf_model* m, int* n) {
    GOPF_API_BEGIN
    if (!m || !n) throw Error("NULL argument");
    *n = (int)m->m.derived.size();
    GOPF_API_END
}

int gopf_model_derived_field_name(gopf_model* m, int index, char* buf, int buf_len) {
    GOPF_API_BEGIN
    if (!m) throw Error("model is NULL");
    if (index < 0 || index >= (int)m->m.derived.size()) throw Error("derived field index out of range");
    copy_name(m->m.derived[index].name, buf, buf_len);
    GOPF_API_END
}

int gopf_model_num_terms(gopf_model* m, int eq, int* n_terms, int* n_denum) {
    GOPF_API_BEGIN
    if (!m || !n_terms || !n_denum) throw Error("NULL argument");
    if (!m->m.initialised) throw Error("Model not initialized");
    if (eq < 0 || eq >= (int)m->m.compiled.size()) throw Error("equation index out of range");
    *n_terms = (int)m->m.compiled[eq].rhs.size();
    *n_denum = (int)m->m.compiled[eq].den.size();
    GOPF_API_END
}

int gopf_model_eq_number(gopf_model* m, const char* field_name, int* eq) {
    GOPF_API_BEGIN
    if (!m || !eq) throw Error("NULL argument");
    *eq = m->m.eq_number(need(field_name, "field_name"));
    GOPF_API_END
}

int gopf_model_destroy(gopf_model* m) {
    GOPF_API_BEGIN
    if (m) {
        if (m->live_solvers > 0) throw Error("gopf_model_destroy: destroy the model's solvers first");
        delete m;
    }
    GOPF_API_END
}

// pf/vandeven.go:13-27
int gopf_vandeven_table(int order, double* out, int n) {
    GOPF_API_BEGIN
    if (!out || n != 1000) throw Error("gopf_vandeven_table: the reference table has exactly 1000 points");
    if (order < 1) throw Error("gopf_vandeven_table: order must be >= 1");
    out[0] = 1.0;
    const double prefactor = std::tgamma((double)(2 * order)) / (std::tgamma((double)order) * std::tgamma((double)order));
    const double dx = 1.0 / 999.0;
    for (int i = 1; i < 1000; ++i) {
        const double x = (double)i * dx;
        const double i2 = std::pow(x * (1 - x), (double)(order - 1));
        const double x1 = x - dx;
        const double i1 = std::pow(x1 * (1.0 - x1), (double)(order - 1));
        out[i] = out[i - 1] - prefactor * 0.5 * (i1 + i2) * dx;
    }
    GOPF_API_END
}

int gopf_solver_create(gopf_model* m, int rank, const int* domain_size, double dt, int device, gopf_solver** out) {
    GOPF_API_BEGIN
    if (!m || !domain_size || !out) throw Error("gopf_solver_create: NULL argument");
    *out = nullptr;
    Solver* s = new Solver(&m->m, rank, domain_size, dt, device);
    gopf_solver* h = new gopf_solver;
    h->s = s;
    h->owner = m;
    m->live_solvers++;
    *out = h;
    GOPF_API_END
}

int gopf_solver_set_stepper(gopf_solver* s, const char* name) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_stepper(need(name, "name"));
    GOPF_API_END
}

int gopf_solver_download_real(gopf_solver* s, int field_index, double* host_out, int big_endian) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->download_real(field_index, host_out, big_endian != 0);
    GOPF_API_END
}

int gopf_solver_set_newton_krylov(gopf_solver* s, int maxiter, double step_size, double tol, int stencil, int restart,
                                  double inner_tol, int max_restarts) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    if (maxiter < 1 || !(step_size > 0.0) || !(tol > 0.0) || (stencil != 2 && stencil != 4 && stencil != 6) || restart < 1 ||
        restart > 200 || !(inner_tol > 0.0) || max_restarts < 1)
        throw Error("gopf_solver_set_newton_krylov: bad option (Stencil is 2, 4 or 6; Restart 1..200)");
    NewtonKrylovOptions o;
    o.maxiter = maxiter;
    o.step_size = step_size;
    o.tol = tol;
    o.stencil = stencil;
    o.restart = restart;
    o.inner_tol = inner_tol;
    o.max_restarts = max_restarts;
    s->s->set_newton_krylov(o);
    GOPF_API_END
}

int gopf_solver_newton_krylov_status(gopf_solver* s, int* converged, int64_t* residual_evaluations) {
    GOPF_API_BEGIN
    if (!s || !converged || !residual_evaluations) throw Error("NULL argument");
    *converged = s->s->last_step_converged() ? 1 : 0;
    *residual_evaluations = (int64_t)s->s->residual_evaluations();
    GOPF_API_END
}

int gopf_solver_set_filter(gopf_solver* s, const double* table, int n) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_filter(table, n);
    GOPF_API_END
}

int gopf_solver_set_stream(gopf_solver* s, void* stream) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_stream(reinterpret_cast<cudaStream_t>(stream));
    GOPF_API_END
}

int gopf_solver_propagate(gopf_solver* s, int nsteps) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->propagate(nsteps);
    GOPF_API_END
}

int gopf_solver_upload(gopf_solver* s) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->upload();
    GOPF_API_END
}

int gopf_solver_step(gopf_solver* s, int nsteps) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->step(nsteps);
    GOPF_API_END
}

int gopf_solver_download(gopf_solver* s) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->download();
    GOPF_API_END
}

int gopf_solver_synchronize(gopf_solver* s) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->synchronize();
    GOPF_API_END
}

int gopf_solver_get_time(gopf_solver* s, double* t) {
    GOPF_API_BEGIN
    if (!s || !t) throw Error("NULL argument");
    *t = s->s->get_time();
    GOPF_API_END
}

int gopf_solver_blocked_layout(gopf_solver* s, int* block_log, int* active) {
    GOPF_API_BEGIN
    if (!s || !block_log || !active) throw Error("gopf_solver_blocked_layout: NULL argument");
    *block_log = s->s->blocked_log();
    *active = s->s->blocked_now() ? 1 : 0;
    GOPF_API_END
}

int gopf_solver_fused_form(gopf_solver* s, int* form, int* derived_form) {
    GOPF_API_BEGIN
    if (!s || !form || !derived_form) throw Error("gopf_solver_fused_form: NULL argument");
    s->s->fused_form(form, derived_form);
    GOPF_API_END
}

int gopf_solver_is_fused(gopf_solver* s, int* fused) {
    GOPF_API_BEGIN
    if (!s || !fused) throw Error("NULL argument");
    *fused = s->s->fused() ? 1 : 0;
    GOPF_API_END
}

int gopf_solver_force_generic(gopf_solver* s, int on) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->force_generic(on != 0);
    GOPF_API_END
}

int gopf_solver_set_jit(gopf_solver* s, int on) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_jit(on != 0);
    GOPF_API_END
}

int gopf_solver_set_jit_inpass(gopf_solver* s, int on) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_jit_inpass(on != 0);
    GOPF_API_END
}

int gopf_solver_jit_kernels(gopf_solver* s, int* count) {
    GOPF_API_BEGIN
    if (!s || !count) throw Error("NULL argument");
    *count = s->s->jit_kernels();
    GOPF_API_END
}

int gopf_solver_jit_log(gopf_solver* s, char* buf, int len) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    copy_name(s->s->jit_log(), buf, len);
    GOPF_API_END
}

int gopf_solver_kernel_launches(gopf_solver* s, int64_t* n, int reset) {
    GOPF_API_BEGIN
    if (!s || !n) throw Error("NULL argument");
    *n = s->s->kernel_launches();
    if (reset) s->s->reset_launch_count();
    GOPF_API_END
}

int gopf_solver_get_spectrum(gopf_solver* s, int index, double* host) {
    GOPF_API_BEGIN
    if (!s || !host) throw Error("NULL argument");
    if (index < 0 || index >= GOPF_MAX_SPECTRA || !s->s->spectrum(index)) throw Error("spectrum not resident on the device");
    s->s->synchronize();
    GOPF_CUDA(cudaMemcpy(host, s->s->spectrum(index), sizeof(cplx) * s->s->plan().N, cudaMemcpyDeviceToHost));
    GOPF_API_END
}

int gopf_solver_lp_multiplier(gopf_solver* s, int slot, double* value) {
    GOPF_API_BEGIN
    if (!s || !value) throw Error("NULL argument");
    *value = s->s->lp_multiplier(slot);
    GOPF_API_END
}

int gopf_solver_sdd_set_orientation(gopf_solver* s, const double* orientation, int64_t len) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->sdd_set_orientation(orientation, (long long)len);
    GOPF_API_END
}

int gopf_solver_sdd_get_orientation(gopf_solver* s, double* host_out) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->sdd_get_orientation(host_out);
    GOPF_API_END
}

int gopf_solver_sdd_set(gopf_solver* s, const char* key, double value) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->sdd_set(need(key, "key"), value);
    GOPF_API_END
}

int gopf_solver_sdd_get(gopf_solver* s, const char* key, double* value) {
    GOPF_API_BEGIN
    if (!s || !value) throw Error("NULL argument");
    *value = s->s->sdd_get(need(key, "key"));
    GOPF_API_END
}

int gopf_solver_term_energy(gopf_solver* s, const char* name, double* energy) {
    GOPF_API_BEGIN
    if (!s || !energy) throw Error("NULL argument");
    *energy = s->s->term_energy(need(name, "name"));
    GOPF_API_END
}

int gopf_solver_download_uint8(gopf_solver* s, int field_index, uint8_t* host_out, double* min_real, double* max_real) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->download_uint8(field_index, host_out, min_real, max_real);
    GOPF_API_END
}

int gopf_solver_charge_current(gopf_solver* s, const char* name, double* host_out) {
    GOPF_API_BEGIN
    if (!s || !host_out) throw Error("NULL argument");
    s->s->charge_current(need(name, "name"), host_out);
    GOPF_API_END
}

int gopf_solver_profile_begin(gopf_solver* s) {
    GOPF_API_BEGIN
    if (!s) throw Error("solver is NULL");
    s->s->set_profiling(true);
    GOPF_API_END
}

int gopf_solver_profile_end(gopf_solver* s, int* n_kernels) {
    GOPF_API_BEGIN
    if (!s || !n_kernels) throw Error("NULL argument");
    s->prof = s->s->collect_profile();
    s->s->set_profiling(false);
    *n_kernels = (int)s->prof.size();
    GOPF_API_END
}

int gopf_solver_profile_get(gopf_solver* s, int i, char* name, int name_len, double* total_ms, int64_t* launches,
                            double* bytes_per_launch) {
    GOPF_API_BEGIN
    if (!s || !total_ms || !launches || !bytes_per_launch) throw Error("NULL argument");
    if (i < 0 || i >= (int)s->prof.size()) throw Error("profile index out of range");
    copy_name(s->prof[i].name, name, name_len);
    *total_ms = s->prof[i].total_ms;
    *launches = s->prof[i].launches;
    *bytes_per_launch = s->prof[i].bytes_per_launch;
    GOPF_API_END
}

int gopf_solver_destroy(gopf_solver* s) {
    GOPF_API_BEGIN
    if (s) {
        delete s->s;
        if (s->owner) s->owner->live_solvers--;
        delete s;
    }
    GOPF_API_END
}

int gopf_host_alloc(int64_t bytes, void** out) {
    GOPF_API_BEGIN
    if (!out || bytes <= 0) throw Error("gopf_host_alloc: bad argument");
    GOPF_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    GOPF_API_END
}

int gopf_host_free(void* p) {
    GOPF_API_BEGIN
    if (p) GOPF_CUDA(cudaFreeHost(p));
    GOPF_API_END
}

}  // extern "C"
