// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for Ipopt's IpTNLP.hpp: just the vocabulary the reference's armtd_NLP (KPR/NLPclass.h) is declared
// with, so that NLPclass.cu compiles unmodified into oracle/_ref/libref_cuda.so.  Ipopt is not installed here; no
// solver exists behind this header — only the callbacks are exercised (oracle/ref_cuda_driver.cu).
#pragma once
namespace Ipopt {
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, STOP_AT_ACCEPTABLE_POINT, LOCAL_INFEASIBILITY, INTERNAL_ERROR };
class IpoptData;
class IpoptCalculatedQuantities;
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
};
}  // namespace Ipopt
