// Host-side Model: the reference's pf.Model (pf/model.go:119-484) with every
// closure replaced by a device-expressible description.  Model::init() is
// pf.Model.Init + pf.Build: it classifies each equation term exactly as
// pf/rhsBuilder.go:26-54 does and compiles the result into a DevKProgram.
#pragma once
#include <complex>
#include <map>
#include <string>
#include <vector>

#include "catalog_terms.cuh"
#include "parser.h"
#include "step_program.h"

namespace gopf {

struct HostField {
    std::string name;
    double* host;  // caller's []complex128 backing array (interleaved), never owned
    size_t n;      // len(Field.Data)
};

enum class DerivedOrigin { Monomial, Function, WhiteNoise, Table };

struct DerivedSpec {
    std::string name;
    DerivedOrigin origin;
    DevDerived dev;
    std::string source;  // monomial description or expression text
    bool used;           // referenced by some compiled term (unused ones are not transformed)
    std::vector<double> table;  // Table: n_steps x N prescribed real values (step s uses row s mod n_steps)
};

enum class UserTermClass { Implicit, Explicit, Mixed };
enum class UserTermKind { SpectralViscosity, PairCorrelation, ExplicitPairCorrelation, IdealMixture, ConservativeNoise, VolumeConservingLP, SquaredGradient, HomogeneousModulusLinElast, TensorialHessian, ChargeTransport };

struct UserTerm {
    std::string name;
    UserTermClass cls;
    UserTermKind kind;
    std::string field;       // PairCorrelation.Field / IdealMixture.Field / VolumeLP.Field / SquaredGradient.Field
    std::string indicator;   // VolumeConservingLP.Indicator
    double prefactor = 1.0;  // Prefactor / Factor
    double c3 = 0.0, c4 = 0.0;  // IdealMixtureTerm.IdealMix (pfc/ideal.go:20-23)
    bool laplacian = false;
    SpectralViscParams sv{};
    PairCorrParams pc{};
    int dim = 0;             // ConservativeNoise.Dim
    std::vector<std::string> current_names;  // ConservativeNoise current fields
    double dt = 0.0;         // VolumeConservingLP.Dt
    int slot = -1;           // index into the DevKProgram special-parameter arrays
    int work_spectrum = -1;  // SquaredGradient / elastic term: spectrum index its result is written to
    double stiffness[81] = {0};  // HomogeneousModulusLinElast.MatProp (elasticity.Rank4.Data, rank4.go:22-24)
    double misfit[9] = {0};      // HomogeneousModulusLinElast.Misfit, row-major 3x3
    TensorHessianParams hessian = {};  // TensorialHessian.K
    // ChargeTransport (chargeTransport.go:29-47): Conductivity(i) tabulated per node, component-major
    // [n_voigt][N]; ExternalField
    std::vector<double> conductivity;
    int n_voigt = 0;
    double external_field[3] = {0.0, 0.0, 0.0};
};

// pf.Source (sourceTerm.go:10-22): the amplitude is a host function of time, evaluated once per
// RHS evaluation and handed to the device as a scalar
typedef double (*SourceFn)(double t, void* user);
struct SourceSpec {
    double pos[3];
    int npos;
    SourceFn fn;
    void* user;
};

struct CompiledEquation {
    std::string text;
    std::string field;
    std::vector<DevTerm> rhs, den;
};

class Model {
public:
    Model();

    size_t N;  // nodes per field (length of Fields[0].Data, model.go:264-269)

    std::vector<HostField> fields;
    std::vector<DerivedSpec> derived;
    std::map<std::string, std::complex<double>> scalars;
    std::vector<std::string> scalar_order;
    std::map<std::string, UserTerm> user_terms;
    std::vector<std::string> equations;  // spaces stripped (model.go:158)
    std::vector<CompiledEquation> compiled;
    std::vector<std::vector<SourceSpec>> sources;  // Model.AllSources (model.go:125, 161)
    std::vector<int> source_spectrum;              // per equation: work spectrum of its sources, -1 if none
    int n_work_spectra = 0;
    // WhiteNoise fields that enter an equation as a plain explicit term are generated directly in
    // k-space (TK_WHITE_NOISE_K, step_program.h) instead of being transformed every step.  Off by default.
    bool kspace_noise = false;
    std::vector<std::pair<int, int>> knoise_slots;  // (parameter slot, derived index) assigned by Init
    bool initialised = false;
    int attached_solvers = 0;  // solvers compiled from this model (their programs do not follow later edits)

    // pf.Model API (model.go:141-162, 322-418)
    void add_field(const std::string& name, size_t n, double* host);
    void add_scalar(const std::string& name, double re, double im);
    void add_equation(const std::string& eq);
    void register_function(const std::string& name, const std::string& expr);  // RegisterFunction, device expression
    void register_white_noise(const std::string& name, double strength, unsigned long long seed);
    void register_table_field(const std::string& name, const double* values, long long n_steps);
    void register_user_term(const UserTerm& t);
    void add_source(int eq_no, const double* pos, int npos, SourceFn f, void* user);  // AddSource (model.go:151-154)
    void register_derived_monomial(const std::string& desc);  // RegisterDerivedField for a monomial description
    void init();                                               // Model.Init (model.go:244-260)

    // queries
    bool is_field_name(const std::string& n) const;    // model.go:211-224 (fields + derived)
    bool is_brick_name(const std::string& n) const;    // model.go:227-234
    bool is_user_term(const std::string& n) const;     // model.go:393-397
    std::vector<std::string> all_field_names() const;  // model.go:198-208
    int field_index(const std::string& n) const;       // among Fields, -1 if absent
    int spectrum_index(const std::string& n) const;    // fields then derived, -1 if absent
    int eq_number(const std::string& field) const;     // model.go:441-455
    int n_spectra() const { return (int)(fields.size() + derived.size()) + n_work_spectra; }

    // compile the k-space program (needs device pointers for filter / multipliers filled by the solver)
    void fill_program(DevKProgram* P, double dt, int rank) const;

private:
    int knoise_base_ = 0;
    void update_derived_fields(const std::string& eq);  // model.go:165-195
    DevTerm concrete_term(const parser::SubStringDelimiter& t) const;  // rhsBuilder.go:125-190
    void apply_prefixes(DevTerm* t, const std::vector<std::string>& prefixes) const;  // rhsBuilder.go:199-242
    CompiledEquation build(const std::string& eq);      // rhsBuilder.go:26-54
    DevDerived compile_monomial(const std::string& desc) const;  // util.go:38-65
    DevDerived compile_expression(const std::string& expr) const;
    void mark_used(int spectrum);
    void panic_on_prefix_in_name(const std::string& name) const;  // model.go:479-484
};

std::complex<double> go_cpow_host(std::complex<double> x, double p);

}  // namespace gopf
