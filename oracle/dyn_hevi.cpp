// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  HEVI: horizontally explicit tendency and the
// vertically implicit column solve.  (Filled in after the HEVE path; see DESIGN.md.)
#include "fe_oracle.hpp"

#include <stdexcept>

namespace feo {
void hevi_numflux_generalvc(const Element&, const Mesh&, const Consts&, const DynState&, vec&) {
  throw std::runtime_error("oracle: HEVI flux not implemented yet");
}
void hevi_cal_tend(const Element&, const Mesh&, const Consts&, const DynState&, double*[5]) {
  throw std::runtime_error("oracle: HEVI tendency not implemented yet");
}
void hevi_cal_vi(const Element&, const Mesh&, const Consts&, const DynState&, const double*[5], double, double, double*[5]) {
  throw std::runtime_error("oracle: HEVI vertical-implicit solve not implemented yet");
}
}  // namespace feo
