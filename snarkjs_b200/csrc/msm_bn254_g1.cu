// MSM instantiation unit: bn254_g1 (coordinate field Fp<BnFq>); the code is msm_group.inl
#define SB_GROUP bn254_g1
#define SB_FIELD Fp<BnFq>
#include "msm_group.inl"
