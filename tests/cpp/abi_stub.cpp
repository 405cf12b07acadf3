// TEST INFRASTRUCTURE.  A recording stand-in for libeuler_b200.so: implements the entry points of include/euler_b200.h
// that the C++ drop-in headers call and stores what it was given, so that the host-side logic of those headers (the
// flattening walk, the parameter keys, the flux gathering, the exception mapping) can be checked without a GPU by
// tests/cpp/flatten_test.cpp.  It computes nothing; it is linked into that test only, never shipped.
#include "abi_stub.hpp"

#include <cstdlib>
#include <cstring>

static StubRecording g_rec;
StubRecording& stub_recording() { return g_rec; }

struct eu_solver { int dummy; };
static eu_solver g_solver;

extern "C" {

int eu_abi_version(void) { return EU_ABI_VERSION; }
const char* eu_last_error(eu_handle) { return g_rec.last_error.c_str(); }

void eu_default_params(eu_params* p)
{
    p->courant_number = 0.5;
    p->method_viscous = p->method_gravity = p->method_capillary = 1;
    p->use_cfl_viscous = p->use_cfl_gravity = p->use_cfl_capillary = 1;
    p->minimum_small_steps = 1; p->maximum_small_steps = 10000;
    p->check_sat = 1; p->clamp_sat = 0;
}

int eu_create(const eu_config* cfg, eu_handle* out)
{
    g_rec.cfg = *cfg;
    ++g_rec.n_create;
    *out = &g_solver;
    return EU_OK;
}
void eu_destroy(eu_handle) { ++g_rec.n_destroy; }
int eu_set_params(eu_handle, const eu_params* p) { g_rec.params = *p; ++g_rec.n_set_params; return EU_OK; }

int eu_grid_begin(eu_handle, int n_global, int n_local, long long n_hf)
{
    g_rec = StubRecording::keepCounters(g_rec);
    g_rec.n_global = n_global; g_rec.n_local = n_local; g_rec.n_hf = n_hf;
    return EU_OK;
}

int eu_grid_append(eu_handle, const eu_grid_chunk* c)
{
    StubRecording& r = g_rec;
    if (c->first_cell != int(r.hf_count.size())) { r.last_error = "chunks out of order"; return EU_ERR_ARG; }
    long long nh = 0;
    for (int i = 0; i < c->n_cells; ++i) nh += c->hf_count[i];
    const long long hf0 = (long long)r.hf_neighbour.size();
    r.hf_count.insert(r.hf_count.end(), c->hf_count, c->hf_count + c->n_cells);
    r.hf_neighbour.insert(r.hf_neighbour.end(), c->hf_neighbour, c->hf_neighbour + nh);
    r.hf_area.insert(r.hf_area.end(), c->hf_area, c->hf_area + nh);
    r.hf_normal.insert(r.hf_normal.end(), c->hf_normal, c->hf_normal + 3*nh);
    r.hf_centroid.insert(r.hf_centroid.end(), c->hf_centroid, c->hf_centroid + 3*nh);
    for (int b = 0; b < c->n_bnd; ++b) {
        r.bnd_hf.push_back(hf0 + c->bnd_hf[b]);
        r.bnd_kind.push_back(c->bnd_kind[b]);
        r.bnd_sat.push_back(c->bnd_sat[b]);
        r.bnd_partner_cell.push_back(c->bnd_partner_cell[b]);
        r.bnd_partner_face.push_back(c->bnd_partner_face[b]);
    }
    r.cell_volume.insert(r.cell_volume.end(), c->cell_volume, c->cell_volume + c->n_cells);
    r.cell_centroid.insert(r.cell_centroid.end(), c->cell_centroid, c->cell_centroid + 3*c->n_cells);
    r.porosity.insert(r.porosity.end(), c->porosity, c->porosity + c->n_cells);
    r.permeability.insert(r.permeability.end(), c->permeability, c->permeability + 9*c->n_cells);
    if (c->rock_id) r.rock_id.insert(r.rock_id.end(), c->rock_id, c->rock_id + c->n_cells);
    ++r.n_chunks;
    return EU_OK;
}

int eu_set_fluid(eu_handle, const eu_fluid* f)
{
    StubRecording& r = g_rec;
    r.fluid = *f;
    r.tab_offset.clear(); r.tab_s.clear();
    for (int k = 0; k < 7; ++k) r.tab_cols[k].clear();
    if (f->n_rocks > 0) {
        r.tab_offset.assign(f->table_offset, f->table_offset + f->n_rocks + 1);
        const int nn = r.tab_offset.back();
        r.tab_s.assign(f->table_s, f->table_s + nn);
        const int ncol = f->mobility_kind == EU_MOB_SCALAR ? 3 : 7;
        for (int k = 0; k < ncol; ++k) r.tab_cols[k].assign(f->table_cols[k], f->table_cols[k] + nn);
    }
    return EU_OK;
}

int eu_grid_end(eu_handle) { g_rec.grid_ended = true; return EU_OK; }
int eu_local_cells(eu_handle) { return g_rec.n_local; }

void* eu_host_alloc(unsigned long long bytes) { ++g_rec.n_host_alloc; return std::malloc(bytes ? bytes : 1); }
void eu_host_free(void* p) { std::free(p); }

int eu_transport_solve(eu_handle, double* saturation, double time, const double gravity[3], const double* hf_flux,
                       int n_src, const int* src_cell, const double* src_rate, eu_report* rep)
{
    StubRecording& r = g_rec;
    r.sat_in.assign(saturation, saturation + r.n_local);
    r.time = time;
    for (int d = 0; d < 3; ++d) r.gravity[d] = gravity[d];
    r.flux.assign(hf_flux, hf_flux + r.n_hf);
    r.src_cell.assign(src_cell, src_cell + n_src);
    r.src_rate.assign(src_rate, src_rate + n_src);
    std::memset(rep, 0, sizeof(*rep));
    rep->status = r.next_status;
    rep->nsteps = 7; rep->attempts = r.next_attempts; rep->bad_cell = r.next_bad_cell; rep->bad_value = r.next_bad_value;
    for (int i = 0; i < r.n_local; ++i) saturation[i] += 1.0;       // a visible, checkable "result"
    if (r.next_status != EU_OK) r.last_error = "stub failure";
    return r.next_status;
}

int eu_compute_residual(eu_handle, const double* saturation, const double gravity[3], const double* hf_flux,
                        int n_src, const int* src_cell, const double* src_rate, int mv, int mg, int mc, double* sat_delta)
{
    StubRecording& r = g_rec;
    r.sat_in.assign(saturation, saturation + r.n_local);
    for (int d = 0; d < 3; ++d) r.gravity[d] = gravity[d];
    if (hf_flux) r.flux.assign(hf_flux, hf_flux + r.n_hf);
    r.src_cell.assign(src_cell, src_cell + n_src);
    r.src_rate.assign(src_rate, src_rate + n_src);
    r.methods[0] = mv; r.methods[1] = mg; r.methods[2] = mc;
    for (int i = 0; i < r.n_local; ++i) sat_delta[i] = double(i);
    return EU_OK;
}

int eu_compute_cap_pressures(eu_handle, const double* saturation, double* pc)
{
    for (int i = 0; i < g_rec.n_local; ++i) pc[i] = 2.0*saturation[i];
    return EU_OK;
}
int eu_cell_velocity(eu_handle, double* v) { for (int i = 0; i < 3*g_rec.n_local; ++i) v[i] = double(i); return EU_OK; }
int eu_phase_velocities(eu_handle, const double*, const double* cv, double* vw, double* vo)
{
    for (int i = 0; i < 3*g_rec.n_local; ++i) { vw[i] = 0.25*cv[i]; vo[i] = 0.75*cv[i]; }
    return EU_OK;
}
int eu_fractional_flow(eu_handle, const double* s, double* f) { for (int i = 0; i < g_rec.n_local; ++i) f[i] = 0.5*s[i]; return EU_OK; }

} // extern "C"

// multi-device plumbing: the CPU tests run the drop-in with one device only; these are never called there
extern "C" {
int eu_device_count(void) { return 1; }
int eu_comm_blob_size(eu_handle) { return 0; }
int eu_comm_export(eu_handle, void*) { return EU_ERR_UNSUPPORTED; }
int eu_comm_connect(eu_handle, int, const void* const*, const int*) { return EU_ERR_UNSUPPORTED; }
int eu_comm_set_allreduce(eu_handle, eu_allreduce_fn, void*) { return EU_ERR_UNSUPPORTED; }
}
