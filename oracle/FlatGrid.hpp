// TEST INFRASTRUCTURE -- not product code.
//
// FlatGrid: a mock model of the GridInterface concept that the reference's explicit
// transport walks (GridInterfaceEuler.hpp:97-456: Face, FaceIterator, Cell,
// CellIterator, GridInterfaceEuler).  The real GridInterfaceEuler<Dune::CpGrid> needs
// Dune + dune-cornerpoint, which are not in this image; the hot path itself is a
// template over the concept, so it runs unmodified over this model.
//
// The grid is a plain CSR of half-faces:
//   cell c owns half-faces [hf_offset[c], hf_offset[c+1]) in its local face order;
//   per half-face: neighbour cell (or -1 on the boundary), boundary id (0 = interior),
//   area, unit outer normal, centroid;  per cell: volume, centroid.
// cell index == iteration order (as for CpGrid, GridInterfaceEuler.hpp:70-80).
#ifndef ORACLE_FLATGRID_HPP
#define ORACLE_FLATGRID_HPP

#include <dune/common/fvector.hh>
#include <climits>
#include <vector>

namespace flatgrid {

    typedef Dune::FieldVector<double, 3> Vec3;

    struct Data {
        int num_cells;
        std::vector<int> hf_offset;      // N+1
        std::vector<int> hf_neighbour;   // H, -1 on boundary
        std::vector<int> hf_bid;         // H, 0 interior
        std::vector<double> hf_area;     // H
        std::vector<double> hf_normal;   // 3H
        std::vector<double> hf_centroid; // 3H
        std::vector<double> cell_volume; // N
        std::vector<double> cell_centroid; // 3N
    };

    inline Vec3 vec3(const double* p) { Vec3 v; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; return v; }

    class CellRef;

    class FaceIterator {
    public:
        typedef Vec3 Vector;
        typedef double Scalar;
        typedef int Index;
        typedef CellRef Cell;
        enum { BoundaryMarkerIndex = -999, LocalEndIndex = INT_MAX };

        FaceIterator() : d_(0), cell_(-1), hf_(-1) {}
        FaceIterator(const Data* d, int cell, int hf) : d_(d), cell_(cell), hf_(hf) {}

        const FaceIterator* operator->() const { return this; }
        const FaceIterator& operator*() const { return *this; }
        FaceIterator& operator++() { ++hf_; return *this; }
        bool operator==(const FaceIterator& o) const { return hf_ == o.hf_; }
        bool operator!=(const FaceIterator& o) const { return hf_ != o.hf_; }
        bool operator<(const FaceIterator& o) const { return hf_ < o.hf_; }

        Scalar area() const { return d_->hf_area[hf_]; }
        Vector centroid() const { return vec3(&d_->hf_centroid[3*hf_]); }
        Vector normal() const { return vec3(&d_->hf_normal[3*hf_]); }
        bool boundary() const { return d_->hf_neighbour[hf_] < 0; }
        int boundaryId() const { return d_->hf_bid[hf_]; }
        inline Cell cell() const;
        Index cellIndex() const { return cell_; }
        inline Cell neighbourCell() const;
        Index neighbourCellIndex() const
        {
            return boundary() ? int(BoundaryMarkerIndex) : d_->hf_neighbour[hf_];
        }
        Index index() const { return hf_; }
        Index localIndex() const { return hf_ - d_->hf_offset[cell_]; }
        Index halfFaceIndex() const { return hf_; }
    private:
        const Data* d_;
        int cell_;
        int hf_;
    };

    class CellRef {
    public:
        typedef flatgrid::FaceIterator FaceIterator;
        typedef Vec3 Vector;
        typedef double Scalar;
        typedef int Index;
        CellRef() : d_(0), c_(-1) {}
        CellRef(const Data* d, int c) : d_(d), c_(c) {}
        FaceIterator facebegin() const { return FaceIterator(d_, c_, d_->hf_offset[c_]); }
        FaceIterator faceend() const { return FaceIterator(d_, c_, d_->hf_offset[c_ + 1]); }
        Scalar volume() const { return d_->cell_volume[c_]; }
        Vector centroid() const { return vec3(&d_->cell_centroid[3*c_]); }
        Index index() const { return c_; }
    protected:
        const Data* d_;
        int c_;
    };

    inline CellRef FaceIterator::cell() const { return CellRef(d_, cell_); }
    inline CellRef FaceIterator::neighbourCell() const { return CellRef(d_, d_->hf_neighbour[hf_]); }

    class CellIterator : public CellRef {
    public:
        CellIterator() {}
        CellIterator(const Data* d, int c) : CellRef(d, c) {}
        const CellIterator* operator->() const { return this; }
        const CellIterator& operator*() const { return *this; }
        CellIterator& operator++() { ++c_; return *this; }
        bool operator==(const CellIterator& o) const { return c_ == o.c_; }
        bool operator!=(const CellIterator& o) const { return c_ != o.c_; }
    };

    class Grid {
    public:
        typedef flatgrid::CellIterator CellIterator;
        typedef Vec3 Vector;
        typedef double Scalar;
        typedef int Index;
        enum { Dimension = 3 };
        Grid() {}
        CellIterator cellbegin() const { return CellIterator(&data_, 0); }
        CellIterator cellend() const { return CellIterator(&data_, data_.num_cells); }
        int numberOfCells() const { return data_.num_cells; }
        int numberOfHalfFaces() const { return int(data_.hf_neighbour.size()); }
        Data& data() { return data_; }
        const Data& data() const { return data_; }
    private:
        Data data_;
    };

} // namespace flatgrid

#endif
