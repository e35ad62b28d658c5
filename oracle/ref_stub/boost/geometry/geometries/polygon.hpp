#include <boost/geometry.hpp>
