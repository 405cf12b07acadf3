// TEST INFRASTRUCTURE: what tests/cpp/abi_stub.cpp recorded (see there).
#ifndef EULER_B200_ABI_STUB_HPP
#define EULER_B200_ABI_STUB_HPP

#include <euler_b200.h>

#include <string>
#include <vector>

struct StubRecording {
    // life cycle
    int n_create = 0, n_destroy = 0, n_set_params = 0, n_chunks = 0, n_host_alloc = 0;
    eu_config cfg;
    eu_params params;
    // grid
    int n_global = 0, n_local = 0;
    long long n_hf = 0;
    bool grid_ended = false;
    std::vector<int> hf_count, hf_neighbour, bnd_kind, bnd_partner_cell, bnd_partner_face, rock_id;
    std::vector<long long> bnd_hf;
    std::vector<double> hf_area, hf_normal, hf_centroid, bnd_sat, cell_volume, cell_centroid, porosity, permeability;
    // fluid
    eu_fluid fluid;
    std::vector<int> tab_offset;
    std::vector<double> tab_s, tab_cols[7];
    // last call
    std::vector<double> sat_in, flux, src_rate;
    std::vector<int> src_cell;
    double time = 0.0, gravity[3] = { 0, 0, 0 };
    int methods[3] = { -1, -1, -1 };
    // what the next eu_transport_solve reports
    int next_status = 0, next_attempts = 1, next_bad_cell = -1;
    double next_bad_value = 0.0;
    std::string last_error;

    static StubRecording keepCounters(const StubRecording& o)
    {
        StubRecording r;
        r.n_create = o.n_create; r.n_destroy = o.n_destroy; r.n_set_params = o.n_set_params; r.n_host_alloc = o.n_host_alloc;
        r.cfg = o.cfg; r.params = o.params;
        r.fluid = o.fluid; r.tab_offset = o.tab_offset; r.tab_s = o.tab_s;
        for (int k = 0; k < 7; ++k) r.tab_cols[k] = o.tab_cols[k];
        return r;
    }
};

StubRecording& stub_recording();

#endif
