// Batched entry points (shared time axis, ragged ensembles, AR(1) surrogates, quantiles).
#include <cuda_runtime.h>

#include "../../include/wwz_b200.h"

extern "C" {

int wwzb200_kirchner_shared_axis(const double *, const double *, int64_t, int64_t, const double *, int64_t,
                                 const double *, int64_t, double, int64_t, int64_t, uint32_t, double *, double *,
                                 double *, double *, double *, double *, double *)
{
    return WWZB200_ERR_ARG;
}
int wwzb200_kirchner_ragged(const double *, const double *, const int64_t *, const double *, const int64_t *,
                            const double *, const int64_t *, const int64_t *, int64_t, double, int64_t, int64_t,
                            uint32_t, double *, double *, double *, double *, double *, double *, double *)
{
    return WWZB200_ERR_ARG;
}
int wwzb200_ar1_surrogates(const double *, int64_t, double, double, double, int64_t, uint64_t, double *)
{
    return WWZB200_ERR_ARG;
}
int wwzb200_quantiles(const double *, int64_t, int64_t, const double *, int64_t, double *) { return WWZB200_ERR_ARG; }
int wwzb200_kirchner_shared_axis_dev(const double *, const double *, int64_t, int64_t, const double *, int64_t,
                                     const double *, int64_t, double, int64_t, int64_t, uint32_t, double *, double *,
                                     double *, double *, double *, double *, double *, void *)
{
    return WWZB200_ERR_ARG;
}
int wwzb200_ar1_surrogates_dev(const double *, int64_t, double, double, double, int64_t, uint64_t, double *, void *)
{
    return WWZB200_ERR_ARG;
}
int wwzb200_quantiles_dev(const double *, int64_t, int64_t, const double *, int64_t, double *, void *)
{
    return WWZB200_ERR_ARG;
}
}
