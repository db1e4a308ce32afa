// C ABI, part 3: slab-sharded step.  Contract: include/gopf_cuda.h.
#include "../../include/gopf_cuda.h"
#include "c_api_types.h"
#include "dist_solver.h"

using namespace gopf;

struct gopf_dist_solver {
    DistSolver* s;
    gopf_model* owner;
};

static cplx* cp(void* p, const char* what) {
    if (!p) throw Error(std::string(what) + " is NULL");
    return reinterpret_cast<cplx*>(p);
}
static const cplx* ccp(const void* p, const char* what) {
    if (!p) throw Error(std::string(what) + " is NULL");
    return reinterpret_cast<const cplx*>(p);
}
static DistSolver& ds(gopf_dist_solver* s) {
    if (!s) throw Error("dist solver is NULL");
    return *s->s;
}

extern "C" {

int gopf_dist_solver_create(gopf_model* m, int n, int world, int rank, double dt, int device, gopf_dist_solver** out) {
    GOPF_API_BEGIN
    if (!m || !out) throw Error("gopf_dist_solver_create: NULL argument");
    *out = nullptr;
    DistSolver* s = new DistSolver(&m->m, n, world, rank, dt, device);
    gopf_dist_solver* h = new gopf_dist_solver;
    h->s = s;
    h->owner = m;
    m->live_solvers++;
    *out = h;
    GOPF_API_END
}

int gopf_dist_solver_set_stream(gopf_dist_solver* s, void* stream) {
    GOPF_API_BEGIN
    ds(s).set_stream(reinterpret_cast<cudaStream_t>(stream));
    GOPF_API_END
}

int gopf_dist_solver_local_cells(gopf_dist_solver* s, int64_t* cells) {
    GOPF_API_BEGIN
    if (!cells) throw Error("cells is NULL");
    *cells = (int64_t)ds(s).local_cells();
    GOPF_API_END
}

int gopf_dist_forward_local(gopf_dist_solver* s, void* w, void* send) {
    GOPF_API_BEGIN
    ds(s).forward_local(cp(w, "w"), cp(send, "send"));
    GOPF_API_END
}

int gopf_dist_forward_finish(gopf_dist_solver* s, void* t) {
    GOPF_API_BEGIN
    ds(s).forward_finish(cp(t, "t"));
    GOPF_API_END
}

int gopf_dist_inverse_start(gopf_dist_solver* s, const void* spectrum, void* t) {
    GOPF_API_BEGIN
    ds(s).inverse_start(ccp(spectrum, "spectrum"), cp(t, "t"));
    GOPF_API_END
}

int gopf_dist_inverse_mid(gopf_dist_solver* s, const void* recv, void* w) {
    GOPF_API_BEGIN
    ds(s).inverse_mid(ccp(recv, "recv"), cp(w, "w"));
    GOPF_API_END
}

int gopf_dist_real_step(gopf_dist_solver* s, void* w) {
    GOPF_API_BEGIN
    ds(s).real_step(cp(w, "w"));
    GOPF_API_END
}

int gopf_dist_forward_mid(gopf_dist_solver* s, const void* w, void* send) {
    GOPF_API_BEGIN
    ds(s).forward_mid(ccp(w, "w"), cp(send, "send"));
    GOPF_API_END
}

int gopf_dist_kspace_step(gopf_dist_solver* s, void* t, void* spectrum) {
    GOPF_API_BEGIN
    ds(s).kspace_step(cp(t, "t"), cp(spectrum, "spectrum"));
    GOPF_API_END
}

int gopf_dist_inverse_finish(gopf_dist_solver* s, void* w, void* real_out) {
    GOPF_API_BEGIN
    ds(s).inverse_finish(cp(w, "w"), cp(real_out, "real_out"));
    GOPF_API_END
}

int gopf_dist_peer_alloc(gopf_dist_solver* s) {
    GOPF_API_BEGIN
    ds(s).peer_alloc();
    GOPF_API_END
}

int gopf_dist_peer_export(gopf_dist_solver* s, int which, void* handle64) {
    GOPF_API_BEGIN
    if (!handle64 || (which != 0 && which != 1)) throw Error("gopf_dist_peer_export: bad argument");
    ds(s).peer_export(which, handle64);
    GOPF_API_END
}

int gopf_dist_peer_import(gopf_dist_solver* s, int which, int rank, const void* handle64) {
    GOPF_API_BEGIN
    if (!handle64 || (which != 0 && which != 1)) throw Error("gopf_dist_peer_import: bad argument");
    ds(s).peer_import(which, rank, handle64);
    GOPF_API_END
}

int gopf_dist_peer_unmap(gopf_dist_solver* s) {
    GOPF_API_BEGIN
    ds(s).peer_unmap();
    GOPF_API_END
}

int gopf_dist_set_grid_cap(gopf_dist_solver* s, int ctas) {
    GOPF_API_BEGIN
    ds(s).set_grid_cap(ctas);
    GOPF_API_END
}

int gopf_dist_peer_local(gopf_dist_solver* s, int which, void** dev_ptr) {
    GOPF_API_BEGIN
    if (!dev_ptr || (which != 0 && which != 1)) throw Error("gopf_dist_peer_local: bad argument");
    *dev_ptr = ds(s).peer_local(which);
    GOPF_API_END
}

int gopf_dist_inverse_start_peer(gopf_dist_solver* s, const void* spectrum) {
    GOPF_API_BEGIN
    ds(s).inverse_start_peer(ccp(spectrum, "spectrum"));
    GOPF_API_END
}

int gopf_dist_forward_mid_peer(gopf_dist_solver* s, const void* w) {
    GOPF_API_BEGIN
    ds(s).forward_mid_peer(ccp(w, "w"));
    GOPF_API_END
}

int gopf_dist_forward_local_peer(gopf_dist_solver* s, void* w) {
    GOPF_API_BEGIN
    ds(s).forward_local_peer(cp(w, "w"));
    GOPF_API_END
}

int gopf_dist_forward_finish_peer(gopf_dist_solver* s, void* spectrum) {
    GOPF_API_BEGIN
    ds(s).forward_finish_peer(cp(spectrum, "spectrum"));
    GOPF_API_END
}

int gopf_dist_kspace_step_peer(gopf_dist_solver* s, void* spectrum) {
    GOPF_API_BEGIN
    ds(s).kspace_step_peer(cp(spectrum, "spectrum"));
    GOPF_API_END
}

static void range_check(int begin, int count, int limit, const char* what) {
    if (begin < 0 || count < 1 || begin + count > limit) throw Error(std::string(what) + ": chunk out of range");
}

int gopf_dist_inverse_mid_planes(gopf_dist_solver* s, const void* recv, void* w, int begin, int count) {
    GOPF_API_BEGIN
    range_check(begin, count, ds(s).slab(), "gopf_dist_inverse_mid_planes");
    ds(s).inverse_mid_planes(ccp(recv, "recv"), cp(w, "w"), begin, count);
    GOPF_API_END
}

int gopf_dist_real_step_planes(gopf_dist_solver* s, void* w, int begin, int count) {
    GOPF_API_BEGIN
    range_check(begin, count, ds(s).slab(), "gopf_dist_real_step_planes");
    ds(s).real_step_planes(cp(w, "w"), begin, count);
    GOPF_API_END
}

int gopf_dist_forward_mid_planes(gopf_dist_solver* s, const void* w, void* send, int begin, int count) {
    GOPF_API_BEGIN
    range_check(begin, count, ds(s).slab(), "gopf_dist_forward_mid_planes");
    ds(s).forward_mid_planes(ccp(w, "w"), cp(send, "send"), begin, count);
    GOPF_API_END
}

int gopf_dist_kspace_step_cols(gopf_dist_solver* s, const void* t_in, void* spectrum, void* t_out, int k1_begin,
                               int k1_count) {
    GOPF_API_BEGIN
    range_check(k1_begin, k1_count, ds(s).slab(), "gopf_dist_kspace_step_cols");
    ds(s).kspace_step_cols(ccp(t_in, "t_in"), cp(spectrum, "spectrum"), cp(t_out, "t_out"), k1_begin, k1_count);
    GOPF_API_END
}

int gopf_dist_exchange_forward(gopf_dist_solver* s, const void* send, int begin, int count) {
    GOPF_API_BEGIN
    range_check(begin, count, ds(s).slab(), "gopf_dist_exchange_forward");
    ds(s).exchange_forward(ccp(send, "send"), begin, count);
    GOPF_API_END
}

int gopf_dist_exchange_inverse(gopf_dist_solver* s, const void* t, int k1_begin, int k1_count) {
    GOPF_API_BEGIN
    range_check(k1_begin, k1_count, ds(s).slab(), "gopf_dist_exchange_inverse");
    ds(s).exchange_inverse(ccp(t, "t"), k1_begin, k1_count);
    GOPF_API_END
}

int gopf_dist_forward_mid_peer_planes(gopf_dist_solver* s, const void* w, int begin, int count, int max_ctas) {
    GOPF_API_BEGIN
    range_check(begin, count, ds(s).slab(), "gopf_dist_forward_mid_peer_planes");
    ds(s).forward_mid_peer_planes(ccp(w, "w"), begin, count, max_ctas);
    GOPF_API_END
}

int gopf_dist_exchange_join(gopf_dist_solver* s) {
    GOPF_API_BEGIN
    ds(s).exchange_join();
    GOPF_API_END
}

int gopf_dist_advance(gopf_dist_solver* s) {
    GOPF_API_BEGIN
    ds(s).advance();
    GOPF_API_END
}

int gopf_dist_solver_get_time(gopf_dist_solver* s, double* t) {
    GOPF_API_BEGIN
    if (!t) throw Error("t is NULL");
    *t = ds(s).get_time();
    GOPF_API_END
}

int gopf_dist_solver_kernel_launches(gopf_dist_solver* s, int64_t* n, int reset) {
    GOPF_API_BEGIN
    if (!n) throw Error("n is NULL");
    *n = ds(s).kernel_launches();
    if (reset) ds(s).reset_launch_count();
    GOPF_API_END
}

int gopf_dist_solver_destroy(gopf_dist_solver* s) {
    GOPF_API_BEGIN
    if (s) {
        delete s->s;
        if (s->owner) s->owner->live_solvers--;
        delete s;
    }
    GOPF_API_END
}

}  // extern "C"
