/* datapath_oracle.c -- CPU restatement of the arithmetic the GPU data-path kernels assume (TEST infrastructure only:
 * imported by tests/, never by the product path).
 *
 * csrc/scene_prepare.cu and csrc/stat_norm.cu reproduce numpy's np.dot on the reference's data path bit for bit by
 * writing every small matrix product as the accumulation order OpenBLAS' sgemm / dgemm kernels use:
 *     out[i][j] = fma(a[i][K-1], b[K-1][j], ... fma(a[i][1], b[1][j], a[i][0] * b[0][j]))
 * (k-sequential FMA chain, first term a plain multiply).  These functions state that order with fmaf / fma so that
 * tests/test_datapath_oracle_cpu.py can pin it against numpy ON THE MACHINE THE TESTS RUN ON: if a different BLAS
 * build ever accumulated differently the CPU suite says so, instead of a GPU parity test failing far from the cause.
 * Reference call sites: pointrcnn/lib/utils/calibration.py:51-71 (float32: lidar_to_rect, rect_to_img),
 * utils/kitti_util.py:141-160 and stat_norm/norm.py:197,218 (float64).  Built with -ffp-contract=off. */
#include <math.h>

/* a (n, K) row-major, b (K, J) row-major -> out (n, J) */
void orc_dot_chain_f32(const float *a, long long n, int K, const float *b, int J, float *out) {
    for (long long i = 0; i < n; ++i)
        for (int j = 0; j < J; ++j) {
            float acc = a[i * K] * b[j];
            for (int k = 1; k < K; ++k) acc = fmaf(a[i * K + k], b[k * J + j], acc);
            out[i * J + j] = acc;
        }
}

void orc_dot_chain_f64(const double *a, long long n, int K, const double *b, int J, double *out) {
    for (long long i = 0; i < n; ++i)
        for (int j = 0; j < J; ++j) {
            double acc = a[i * K] * b[j];
            for (int k = 1; k < K; ++k) acc = fma(a[i * K + k], b[k * J + j], acc);
            out[i * J + j] = acc;
        }
}
