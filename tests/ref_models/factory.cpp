/*
 * factory.cpp -- TEST INFRASTRUCTURE.  Hands out model objects whose classes come from the REFERENCE's own, unchanged source files
 * (compiled next to this file by build_ref_models.py against ntrtsim_b200/host/compat), so that a test can run them through the
 * B200 host API and the device: the boundary claim "models are built on the host in C++ exactly as today".
 * The headers below are the reference's (found through -I /root/reference/src/...), not this repo's tgb_examples.h.
 */
#include "PrismModel.h"                 /* src/examples/3_prism */
#include "T6Model.h"                    /* src/examples/SUPERball */
#include "EscapeModel.h"                /* src/examples/craterEscape */
#include "tsTestRig.h"                  /* src/examples/motorModel */
#include "TetraSpineLearningModel.h"    /* src/examples/learningSpines/TetraSpine */
#include "FlemonsSpineModelLearning.h"  /* src/examples/learningSpines/TetrahedralComplex */
#include "tgCraterDeep.h"               /* src/models/obstacles */
#include "tgCraterShallow.h"
#include "tgStairs.h"

#include <string>

extern "C" tgModel* tgb_make_model(const char* which_c) {
    const std::string which(which_c ? which_c : "");
    if (which == "3_prism" || which == "3_prism_noground") return new PrismModel();
    if (which == "SUPERball" || which == "SUPERball_flat" || which == "SUPERball_stairs") return new T6Model();
    if (which == "Escape" || which == "Escape_c3" || which == "Escape_crater" || which == "Escape_crater_shallow") return new EscapeModel();
    if (which == "Escape_crater:obstacle") return new tgCraterDeep(btVector3(0, 0, 0));
    if (which == "Escape_crater_shallow:obstacle") return new tgCraterShallow(btVector3(0, 0, 0));
    if (which == "SUPERball_stairs:obstacle") {
        tgStairs::Config c(btVector3(0.0, 2.0, -10.0), 0.5, 0.0, 6, 20.0, 5.0, 1.0, 0.0);
        return new tgStairs(c);
    }
    if (which == "tsTestRig_basic") return new tsTestRig(false);
    if (which == "tsTestRig_kinematic") return new tsTestRig(true);
    if (which == "TetraSpine") return new TetraSpineLearningModel(3);
    if (which == "FlemonsSpine") return new FlemonsSpineModelLearning(12);
    if (which == "FlemonsSpine3") return new FlemonsSpineModelLearning(3);
    return 0;
}
