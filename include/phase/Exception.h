// Exception.h -- same class and message format as the reference
// (src/System/Exception.h:7-18, Exception.cpp:3-6); every non-zero status of the
// C ABI is converted into one of these by phase::check().
#ifndef PHASE_B200_EXCEPTION_H
#define PHASE_B200_EXCEPTION_H
#include <exception>
#include <string>

#include "../phase_b200.h"

class Exception : public std::exception {
public:
  explicit Exception(const std::string &className, const std::string &methodName,
                     const std::string &description)
      : message_(className + "::" + methodName + " -> " + description) {}
  virtual ~Exception() throw() {}
  virtual const char *what() const throw() { return message_.c_str(); }

protected:
  std::string message_;
};

namespace phase {
inline int check(int rc, const char *cls, const char *method) {
  if (rc < 0) throw Exception(cls, method, phb_last_error());
  return rc;
}
}  // namespace phase
#endif
