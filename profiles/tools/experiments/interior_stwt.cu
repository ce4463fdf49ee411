#define MSB_ST(p, v) __stwt(p, v)
#include "interior_plain_st.cu"
