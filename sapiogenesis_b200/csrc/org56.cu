#include "org_tu.cuh"
ORG_TU(56)
