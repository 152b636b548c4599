// OPR_Fourier_X_Forward / _Backward and OPR_Fourier_Z_Forward / _Backward of the reference
// (src/operators/opr_fourier.f90:219-433; plans :54-208) as stand-alone entry points.
//
// The reference plans FFTW3 real-to-complex transforms along x for ny*nz lines (each z-plane holds ny lines of nx/2+1
// complex numbers, the last one being the Nyquist mode) and complex transforms along z with stride (nx/2+1)*ny;
// forward = FFTW_FORWARD (exponent -i), backward = FFTW_BACKWARD, both unnormalised (OPR_Poisson divides by nx*nz itself,
// opr_elliptic.f90:292).  north_star asks for cuFFT here: the same geometry as cufftPlanMany D2Z / Z2D / Z2Z plans, cached
// per grid.  The Poisson solver keeps its own plans (poisson.cu); these are for hosts that call the transforms directly
// (spectra, OPR_Fourier_F/B-style filters).
#include "../../include/tlab_gpu.h"
#include "context.h"
#include "trp.h"
#include <cufft.h>
#include <map>
#include <tuple>

namespace tlab {
namespace {

struct FourierPlans {
    cufftHandle fx = 0, bx = 0, z = 0;
    bool has_z = false;
};

std::map<std::tuple<int, int, int>, FourierPlans>& plan_cache() {
    static std::map<std::tuple<int, int, int>, FourierPlans> c;
    return c;
}

int cufft_fail(cufftResult r, const char* what) {
    return fail(TLAB_ERR_CUDA, std::string(what) + ": cuFFT error " + std::to_string((int)r));
}

int get_plans(int nx, int ny, int nz, FourierPlans** out) {
    if (nx < 2 || nx % 2 != 0 || ny < 1 || nz < 1) return fail(TLAB_ERR_DIMGRID, "OPR_Fourier: nx must be even, ny, nz >= 1");
    auto& cache = plan_cache();
    const auto key = std::make_tuple(nx, ny, nz);
    auto it = cache.find(key);
    if (it == cache.end()) {
        FourierPlans p;
        const int nxh = nx / 2 + 1;
        int n1[1] = {nx};
        cufftResult r = cufftPlanMany(&p.fx, 1, n1, n1, 1, nx, n1, 1, nxh, CUFFT_D2Z, ny * nz);
        if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftPlanMany D2Z");
        r = cufftPlanMany(&p.bx, 1, n1, n1, 1, nxh, n1, 1, nx, CUFFT_Z2D, ny * nz);
        if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftPlanMany Z2D");
        if (nz > 1) {
            int n3[1] = {nz};
            const int howmany = nxh * ny;
            r = cufftPlanMany(&p.z, 1, n3, n3, howmany, 1, n3, howmany, 1, CUFFT_Z2Z, howmany);
            if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftPlanMany Z2Z");
            p.has_z = true;
        }
        if (cache.size() >= 8) {           // a host works on one or two grids; do not hoard work areas
            for (auto& kv : cache) { cufftDestroy(kv.second.fx); cufftDestroy(kv.second.bx); if (kv.second.has_z) cufftDestroy(kv.second.z); }
            cache.clear();
        }
        it = cache.emplace(key, p).first;
    }
    cufftSetStream(it->second.fx, ctx().stream);
    cufftSetStream(it->second.bx, ctx().stream);
    if (it->second.has_z) cufftSetStream(it->second.z, ctx().stream);
    *out = &it->second;
    return 0;
}

int ready_single_domain() {
    if (!ctx().ready) { if (int rc = tlab_gpu_init(-1)) return rc; }
    if (trp().P > 1) return fail(TLAB_ERR_UNDEVELOP, "OPR_Fourier entry points: single-domain only (the split-domain transforms live inside OPR_Poisson)");
    return 0;
}

}  // namespace

void fourier_release() {
    for (auto& kv : plan_cache()) { cufftDestroy(kv.second.fx); cufftDestroy(kv.second.bx); if (kv.second.has_z) cufftDestroy(kv.second.z); }
    plan_cache().clear();
}

}  // namespace tlab

using namespace tlab;

extern "C" {

int tlab_opr_fourier_x_forward(int nx, int ny, int nz, const double* in, double* out) {
    if (int rc = ready_single_domain()) return rc;
    if (!in || !out) return fail(TLAB_ERR_OPTION, "OPR_Fourier_X_Forward: null array");
    FourierPlans* p;
    if (int rc = get_plans(nx, ny, nz, &p)) return rc;
    ProfScope ps(PC_FFT);
    const cufftResult r = cufftExecD2Z(p->fx, const_cast<double*>(in), reinterpret_cast<cufftDoubleComplex*>(out));
    if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftExecD2Z");
    return finish();
}

int tlab_opr_fourier_x_backward(int nx, int ny, int nz, double* in, double* out) {
    if (int rc = ready_single_domain()) return rc;
    if (!in || !out) return fail(TLAB_ERR_OPTION, "OPR_Fourier_X_Backward: null array");
    FourierPlans* p;
    if (int rc = get_plans(nx, ny, nz, &p)) return rc;
    ProfScope ps(PC_FFT);
    const cufftResult r = cufftExecZ2D(p->bx, reinterpret_cast<cufftDoubleComplex*>(in), out);
    if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftExecZ2D");
    return finish();
}

static int fourier_z(int nx, int ny, int nz, double* in, double* out, int dir) {
    if (int rc = ready_single_domain()) return rc;
    if (!in || !out) return fail(TLAB_ERR_OPTION, "OPR_Fourier_Z: null array");
    FourierPlans* p;
    if (int rc = get_plans(nx, ny, nz, &p)) return rc;
    ProfScope ps(PC_FFT);
    if (!p->has_z) {          // 2-D case: the transform along a single plane is the identity (opr_fourier.f90:372-378)
        if (in != out) cudaMemcpyAsync(out, in, (size_t)(nx / 2 + 1) * ny * 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream);
        return finish();
    }
    const cufftResult r = cufftExecZ2Z(p->z, reinterpret_cast<cufftDoubleComplex*>(in), reinterpret_cast<cufftDoubleComplex*>(out), dir);
    if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftExecZ2Z");
    return finish();
}

int tlab_opr_fourier_z_forward(int nx, int ny, int nz, double* in, double* out) { return fourier_z(nx, ny, nz, in, out, CUFFT_FORWARD); }
int tlab_opr_fourier_z_backward(int nx, int ny, int nz, double* in, double* out) { return fourier_z(nx, ny, nz, in, out, CUFFT_INVERSE); }

}  // extern "C"
