#define MSB_ST(p, v) __stcg(p, v)
#include "interior_plain_st.cu"
