// Stand-in for <boost/property_tree/ptree.hpp> (Boost is not installed here).
// Only used to compile the reference's src/Math/*.cpp IN PLACE for oracle/_ref;
// those files never touch a ptree beyond naming the type in
// SparseMatrixSolver::setup's signature.
#ifndef PHASE_ORACLE_PTREE_STUB
#define PHASE_ORACLE_PTREE_STUB
#include <algorithm>
#include <memory>
#include <ostream>
#include <string>
namespace boost { namespace property_tree { class ptree {}; } }
#endif
