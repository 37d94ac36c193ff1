// A plug-in module for path-planner-cli --plugins: registers a Planner class under a new name, as a user library does
// with MRPT's class registry in the reference (src/registerClasses.cpp:16-37; path-planner-cli.cpp:452-461).
#include <iostream>

#include "mpp.h"

namespace testplugin
{
class LoggingPlanner : public mpp::TPS_Astar
{
   public:
    using mpp::TPS_Astar::TPS_Astar;
    mpp::PlannerOutput plan(const mpp::PlannerInput& in) override
    {
        std::cout << "[plugin] testplugin::LoggingPlanner::plan" << std::endl;
        return mpp::TPS_Astar::plan(in);
    }
};
}  // namespace testplugin
using testplugin::LoggingPlanner;
MPP_REGISTER_PLANNER("testplugin::LoggingPlanner", LoggingPlanner)
