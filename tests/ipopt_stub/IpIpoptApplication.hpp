// TEST INFRASTRUCTURE — see IpTNLP.hpp in this directory.  IpoptApplication with the option setters and return statuses the
// reference's main() uses (KPR/armour_main.cu:254-317).  OptimizeTNLP runs the repo's stand-in Gauss-Newton solver through the
// TNLP virtual interface.  Environment hooks for the tests: IPOPT_STUB_STATUS = Maximum_CpuTime_Exceeded | Invalid_Option |
// Initialize_Failure makes the stub report that outcome (the CPU-time case still solves first, like a time-out with a usable iterate).
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "IpTNLP.hpp"
#include "standin_solver.hpp"   // armour-dev_b200/host: the solver is a template over the TNLP type

namespace Ipopt {
enum ApplicationReturnStatus { Solve_Succeeded = 0, Solved_To_Acceptable_Level = 1, Infeasible_Problem_Detected = 2, Search_Direction_Becomes_Too_Small = 3,
                               Maximum_Iterations_Exceeded = -1, Restoration_Failed = -2, Error_In_Step_Computation = -3, Maximum_CpuTime_Exceeded = -4,
                               Maximum_WallTime_Exceeded = -5, Not_Enough_Degrees_Of_Freedom = -10, Invalid_Problem_Definition = -11, Invalid_Option = -12,
                               Invalid_Number_Detected = -13, Unrecoverable_Exception = -100, NonIpopt_Exception_Thrown = -101, Insufficient_Memory = -102,
                               Internal_Error = -199 };
class OptionsList : public ReferencedObject {
public:
    bool SetNumericValue(const std::string& k, Number v) { num[k] = v; return true; }
    bool SetIntegerValue(const std::string& k, Index v) { integer[k] = v; return true; }
    bool SetStringValue(const std::string& k, const std::string& v) { str[k] = v; return true; }
    std::map<std::string, Number> num;
    std::map<std::string, Index> integer;
    std::map<std::string, std::string> str;
};
class IpoptApplication : public ReferencedObject {
public:
    IpoptApplication() : options_(new OptionsList()) {}
    SmartPtr<OptionsList> Options() { return options_; }
    ApplicationReturnStatus Initialize() {
        const char* s = getenv("IPOPT_STUB_STATUS");
        return (s && !strcmp(s, "Initialize_Failure")) ? Invalid_Option : Solve_Succeeded;
    }
    ApplicationReturnStatus OptimizeTNLP(const SmartPtr<TNLP>& tnlp) {
        const char* s = getenv("IPOPT_STUB_STATUS");
        if (s && !strcmp(s, "Invalid_Option")) return Invalid_Option;   // what Ipopt reports when the HSL library is missing: no solve
        double x[64];
        StandinResult r = standin_solve(*tnlp, x);
        if (s && !strcmp(s, "Maximum_CpuTime_Exceeded")) return Maximum_CpuTime_Exceeded;
        return r.converged ? Solve_Succeeded : Maximum_Iterations_Exceeded;
    }
private:
    SmartPtr<OptionsList> options_;
};
inline IpoptApplication* IpoptApplicationFactory() { return new IpoptApplication(); }
}  // namespace Ipopt
