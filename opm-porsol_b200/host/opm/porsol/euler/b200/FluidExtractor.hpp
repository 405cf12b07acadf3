// Extraction of the fluid / rock description from the reference's property classes WITHOUT modifying them.
//
// The hot path only ever *evaluates* the rock curves (RockJfunc::krw/kro/capPress, RockJfunc.hpp:70-112), but
// the device needs the table nodes.  They are private members (RockJfunc.hpp:220-228,
// RockAnisotropicRelperm.hpp:154-159, ReservoirPropertyCommon.hpp:246-247, opm-core
// NonuniformTableLinear::x_values_/y_values_).  Explicit template instantiation may name private members
// ([temp.spec]/6), which gives read access with zero changes to opm-porsol or opm-core.  A maintainer who
// prefers accessors can add them and specialise FluidExtractor instead.
#ifndef OPM_B200_FLUIDEXTRACTOR_HEADER
#define OPM_B200_FLUIDEXTRACTOR_HEADER

#include <opm/porsol/common/ReservoirPropertyCapillary.hpp>
#include <opm/porsol/common/ReservoirPropertyCapillaryAnisotropicRelperm.hpp>

#include <euler_b200.h>

#include <stdexcept>
#include <vector>

namespace Opm {
namespace b200 {

    struct FluidDescription {
        int mobility_kind;
        double viscosity[2], density[2], cfl_factor[3];
        int use_j;
        double sigma_cos_theta;
        int n_rocks;
        std::vector<int> table_offset;
        std::vector<double> table_s;
        std::vector<double> cols[7];
        const std::vector<int>* cell_to_rock;
        FluidDescription() : mobility_kind(EU_MOB_SCALAR), use_j(0), sigma_cos_theta(1.0), n_rocks(0), cell_to_rock(0) {}
        int rockOfCell(int c) const { return (n_rocks > 0 && cell_to_rock) ? (*cell_to_rock)[c] : 0; }
        void fill(eu_fluid& f) const
        {
            f.mobility_kind = mobility_kind;
            for (int i = 0; i < 2; ++i) { f.viscosity[i] = viscosity[i]; f.density[i] = density[i]; }
            for (int i = 0; i < 3; ++i) f.cfl_factor[i] = cfl_factor[i];
            f.use_jfunction_scaling = use_j;
            f.sigma_cos_theta = sigma_cos_theta;
            f.n_rocks = n_rocks;
            f.table_offset = table_offset.data();
            f.table_s = table_s.data();
            for (int k = 0; k < 7; ++k) f.table_cols[k] = cols[k].empty() ? table_s.data() : cols[k].data();
        }
    };

    namespace access {
        // the standard-conforming "explicit instantiation" private-member accessor
        template <class Tag, typename Tag::type Member>
        struct Rob { friend typename Tag::type get(Tag) { return Member; } };

        typedef NonuniformTableLinear<double> Tab;
#define EULER_B200_MEMBER(Tag, Class, Type, member)                 \
        struct Tag { typedef Type Class::*type; friend type get(Tag); }; \
        template struct Rob<Tag, &Class::member>
        EULER_B200_MEMBER(TabX, Tab, std::vector<double>, x_values_);
        EULER_B200_MEMBER(TabY, Tab, std::vector<double>, y_values_);
        EULER_B200_MEMBER(JKrw, RockJfunc, Tab, krw_);
        EULER_B200_MEMBER(JKro, RockJfunc, Tab, kro_);
        EULER_B200_MEMBER(JJ, RockJfunc, Tab, Jfunc_);
        EULER_B200_MEMBER(JUseJ, RockJfunc, bool, use_jfunction_scaling_);
        EULER_B200_MEMBER(JSigma, RockJfunc, double, sigma_cos_theta_);
        EULER_B200_MEMBER(APc, RockAnisotropicRelperm, Tab, cap_press_);
        typedef Tab Tab2[2];
        EULER_B200_MEMBER(AKx, RockAnisotropicRelperm, Tab2, krxx_);
        EULER_B200_MEMBER(AKy, RockAnisotropicRelperm, Tab2, kryy_);
        EULER_B200_MEMBER(AKz, RockAnisotropicRelperm, Tab2, krzz_);
        typedef ReservoirPropertyCommon<3, ReservoirPropertyCapillary<3>, RockJfunc> CommonJ;
        typedef ReservoirPropertyCommon<3, ReservoirPropertyCapillaryAnisotropicRelperm<3>, RockAnisotropicRelperm> CommonA;
        EULER_B200_MEMBER(CJRock, CommonJ, std::vector<RockJfunc>, rock_);
        EULER_B200_MEMBER(CJCell, CommonJ, std::vector<int>, cell_to_rock_);
        EULER_B200_MEMBER(CARock, CommonA, std::vector<RockAnisotropicRelperm>, rock_);
        EULER_B200_MEMBER(CACell, CommonA, std::vector<int>, cell_to_rock_);
#undef EULER_B200_MEMBER

        inline const std::vector<double>& xs(const Tab& t) { return t.*get(TabX()); }
        inline const std::vector<double>& ys(const Tab& t) { return t.*get(TabY()); }
    }

    template <class RP>
    struct FluidExtractor;     // customisation point: specialise for other property classes

    template <class RP>
    inline void extractCommon(const RP& rp, FluidDescription& fd)
    {
        fd.viscosity[0] = rp.viscosityFirstPhase();
        fd.viscosity[1] = rp.viscositySecondPhase();
        fd.density[0] = rp.densityFirstPhase();
        fd.density[1] = rp.densitySecondPhase();
        fd.cfl_factor[0] = rp.cflFactor();
        fd.cfl_factor[1] = rp.cflFactorGravity();
        fd.cfl_factor[2] = rp.cflFactorCapillary();
    }

    template <>
    struct FluidExtractor<ReservoirPropertyCapillary<3> > {
        static void extract(const ReservoirPropertyCapillary<3>& rp, int /*num_cells*/, FluidDescription& fd)
        {
            using namespace access;
            fd.mobility_kind = EU_MOB_SCALAR;
            extractCommon(rp, fd);
            const CommonJ& base = rp;
            const std::vector<RockJfunc>& rocks = base.*get(CJRock());
            fd.cell_to_rock = &(base.*get(CJCell()));
            fd.n_rocks = int(rocks.size());
            fd.table_offset.assign(1, 0);
            for (size_t r = 0; r < rocks.size(); ++r) {
                const Tab& krw = rocks[r].*get(JKrw());
                const Tab& kro = rocks[r].*get(JKro());
                const Tab& J = rocks[r].*get(JJ());
                // one saturation column per rock in the C ABI: the three curves must share their nodes (they do when read
                // from the reference's rock files, RockJfunc.hpp:196-228, one row per saturation)
                if (xs(kro) != xs(krw) || xs(J) != xs(krw))
                    throw std::runtime_error("rock table: krw, kro and J must share their saturation nodes");
                if (ys(krw).size() != xs(krw).size() || ys(kro).size() != xs(krw).size() || ys(J).size() != xs(krw).size())
                    throw std::runtime_error("rock table: columns of different lengths");
                fd.table_s.insert(fd.table_s.end(), xs(krw).begin(), xs(krw).end());
                fd.cols[0].insert(fd.cols[0].end(), ys(krw).begin(), ys(krw).end());
                fd.cols[1].insert(fd.cols[1].end(), ys(kro).begin(), ys(kro).end());
                fd.cols[2].insert(fd.cols[2].end(), ys(J).begin(), ys(J).end());
                fd.table_offset.push_back(int(fd.table_s.size()));
                fd.use_j = (rocks[r].*get(JUseJ())) ? 1 : 0;
                fd.sigma_cos_theta = rocks[r].*get(JSigma());
            }
            if (rocks.empty()) fd.table_s.assign(1, 0.0);
        }
    };

    template <>
    struct FluidExtractor<ReservoirPropertyCapillaryAnisotropicRelperm<3> > {
        static void extract(const ReservoirPropertyCapillaryAnisotropicRelperm<3>& rp, int, FluidDescription& fd)
        {
            using namespace access;
            fd.mobility_kind = EU_MOB_DIAGONAL;
            extractCommon(rp, fd);
            const CommonA& base = rp;
            const std::vector<RockAnisotropicRelperm>& rocks = base.*get(CARock());
            fd.cell_to_rock = &(base.*get(CACell()));
            fd.n_rocks = int(rocks.size());
            fd.table_offset.assign(1, 0);
            for (size_t r = 0; r < rocks.size(); ++r) {
                const Tab& pc = rocks[r].*get(APc());
                fd.table_s.insert(fd.table_s.end(), xs(pc).begin(), xs(pc).end());
                fd.cols[0].insert(fd.cols[0].end(), ys(pc).begin(), ys(pc).end());
                for (int ph = 0; ph < 2; ++ph) {
                    const Tab& kx = (rocks[r].*get(AKx()))[ph];
                    const Tab& ky = (rocks[r].*get(AKy()))[ph];
                    const Tab& kz = (rocks[r].*get(AKz()))[ph];
                    if (xs(kx) != xs(pc) || xs(ky) != xs(pc) || xs(kz) != xs(pc))
                        throw std::runtime_error("anisotropic rock: phase tables must share saturation nodes");
                    fd.cols[1 + 3*ph].insert(fd.cols[1 + 3*ph].end(), ys(kx).begin(), ys(kx).end());
                    fd.cols[2 + 3*ph].insert(fd.cols[2 + 3*ph].end(), ys(ky).begin(), ys(ky).end());
                    fd.cols[3 + 3*ph].insert(fd.cols[3 + 3*ph].end(), ys(kz).begin(), ys(kz).end());
                }
                fd.table_offset.push_back(int(fd.table_s.size()));
            }
            if (rocks.empty()) fd.table_s.assign(1, 0.0);
        }
    };

} // namespace b200
} // namespace Opm

#endif
