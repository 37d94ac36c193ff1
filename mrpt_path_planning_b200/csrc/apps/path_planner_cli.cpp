// path-planner-cli — one-shot planner front end with the reference tool's flags and printed
// output (apps/path-planner-cli/path-planner-cli.cpp:30-148 flags, :200-430 do_plan_path), running
// headless on the GPU library: mpp::TPS_Astar evaluates collisions and evaluator costs on the B200.
//
// Differences on purpose: no GUI (the reference always opens a window at the end; here the
// result is text and the optional CSV only), obstacle inputs are text point lists only
// (.txt/.pts; grid maps and images need MRPT's occupancy grid, which is out of the hot path),
// `--planner` accepts mpp::TPS_Astar (and the alias mpp::TPS_Astar_B200) since class registration
// by name goes through MRPT's RTTI, and `--device N` picks the CUDA device.
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>

#include "../../../include/mppb.h"
#include "../host/mpp.h"

namespace
{
struct Args
{
    std::map<std::string, std::string> values;
    std::set<std::string>              switches;
    bool                               isSet(const std::string& k) const { return values.count(k) || switches.count(k); }
    std::string get(const std::string& k, const std::string& def = "") const
    {
        const auto it = values.find(k);
        return it == values.end() ? def : it->second;
    }
};

const std::map<std::string, std::string> kShort = {{"-o", "--obstacles"},   {"-p", "--planner"},    {"-v", "--verbose"},
                                                   {"-c", "--ptg-config"},  {"-s", "--start-pose"}, {"-g", "--goal-pose"}};
const std::set<std::string> kValueFlags = {"--obstacles", "--planner", "--verbose", "--obstacles-gridimage-resolution",
                                           "--interpolarion-period", "--ptg-config", "--planner-parameters",
                                           "--write-planner-parameters", "--config-section", "--start-pose", "--start-vel",
                                           "--goal-pose", "--bbox-margin", "--goal-vel", "--random-seed", "--plugins",
                                           "--costmap-obstacles", "--waypoints", "--waypoints-parameters",
                                           "--save-interpolated-path", "--device"};
const std::set<std::string> kSwitchFlags = {"--show-tree", "--ignore-obstacles-bbox", "--no-refine", "--show-edge-weights",
                                            "--print-path-edges", "--play-animation", "--no-gui", "--help", "-h"};

Args parse(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; i++)
    {
        std::string f = argv[i];
        if (kShort.count(f)) f = kShort.at(f);
        std::string inlineValue;
        const auto  eq       = f.find('=');
        bool        hasInline = false;
        if (f.rfind("--", 0) == 0 && eq != std::string::npos)
        {
            inlineValue = f.substr(eq + 1);
            f           = f.substr(0, eq);
            hasInline   = true;
        }
        if (kValueFlags.count(f))
        {
            if (hasInline)
                a.values[f] = inlineValue;
            else
            {
                if (i + 1 >= argc) throw std::runtime_error("PARSE ERROR: missing value for argument " + f);
                a.values[f] = argv[++i];
            }
        }
        else if (kSwitchFlags.count(f))
            a.switches.insert(f);
        else
            throw std::runtime_error("PARSE ERROR: unknown argument " + f);
    }
    return a;
}

mpp::TTwist2D twist_from_string(const std::string& s)
{
    const auto p = mpp::PoseOrPoint::FromString(s);  // "[vx vy omega_deg]"
    if (!p.isPose()) throw std::runtime_error("Velocity must be \"[vx vy omega_deg]\"");
    return mpp::TTwist2D(p.pose().x, p.pose().y, p.pose().phi);
}

void usage()
{
    std::cout << "USAGE: path-planner-cli -o obs.txt -c ptgs.ini -s \"[x y phi_deg]\" -g \"[x y phi_deg]\"|\"[x y]\"\n"
                 "  [-p mpp::TPS_Astar] [--planner-parameters tps-astar.yaml] [--write-planner-parameters out.yaml]\n"
                 "  [--config-section SelfDriving] [--start-vel \"[vx vy w_deg]\"] [--goal-vel \"[vx vy w_deg]\"]\n"
                 "  [--bbox-margin 2.0] [--ignore-obstacles-bbox] [--costmap-obstacles costmap.yaml]\n"
                 "  [--waypoints wps.yaml] [--waypoints-parameters params.yaml] [--no-refine] [--print-path-edges]\n"
                 "  [--save-interpolated-path path.csv] [--interpolarion-period 0.25] [--device 0]\n";
}

void do_plan_path(const Args& a)
{
    // obstacles
    const std::string obsFile = a.get("--obstacles");
    auto              obsPts  = std::make_shared<mpp::PointCloud>();
    {
        std::ifstream probe(obsFile);
        if (!probe.is_open()) throw std::runtime_error("Assert file existence failed: " + obsFile);
        const auto dot = obsFile.rfind('.');
        std::string ext = dot == std::string::npos ? "" : obsFile.substr(dot + 1);
        for (auto& c : ext) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
        if (ext != "txt" && ext != "pts")
            throw std::runtime_error("Unsupported obstacle file type '." + ext + "' (this build reads .txt/.pts point lists)");
        if (!obsPts->load2D_from_text_file(obsFile))
            throw std::runtime_error("Cannot read obstacle point cloud from: `" + obsFile + "`");
    }
    auto obs = mpp::ObstacleSource::FromStaticPointcloud(obsPts);

    mpp::PlannerInput pi;
    {
        const auto s = mpp::PoseOrPoint::FromString(a.get("--start-pose"));
        if (!s.isPose()) throw std::runtime_error("Start pose must be \"[x y phi_deg]\"");
        pi.stateStart.pose = s.pose();
    }
    if (a.isSet("--start-vel")) pi.stateStart.vel = twist_from_string(a.get("--start-vel"));
    pi.stateGoal.state = mpp::PoseOrPoint::FromString(a.get("--goal-pose"));
    if (a.isSet("--goal-vel")) pi.stateGoal.vel = twist_from_string(a.get("--goal-vel"));
    pi.obstacles.push_back(obs);

    // world bounding box in float, as TBoundingBoxf (path-planner-cli.cpp:219-247)
    float minx, miny, maxx, maxy;
    if (!a.isSet("--ignore-obstacles-bbox"))
        obsPts->boundingBox(minx, miny, maxx, maxy);
    else
    {
        minx = miny = std::numeric_limits<float>::max();
        maxx = maxy = -std::numeric_limits<float>::max();
    }
    {
        const float m    = static_cast<float>(std::stod(a.get("--bbox-margin", "2.0")));
        const auto  goal = pi.stateGoal.asSE2KinState().pose;
        const float px[2] = {static_cast<float>(pi.stateStart.pose.x), static_cast<float>(goal.x)};
        const float py[2] = {static_cast<float>(pi.stateStart.pose.y), static_cast<float>(goal.y)};
        for (int i = 0; i < 2; i++)
            for (const float sgn : {-1.f, 1.f})
            {
                minx = std::min(minx, px[i] + sgn * m);
                maxx = std::max(maxx, px[i] + sgn * m);
                miny = std::min(miny, py[i] + sgn * m);
                maxy = std::max(maxy, py[i] + sgn * m);
            }
    }
    pi.worldBboxMax = mpp::TPose2D(maxx, maxy, M_PI);
    pi.worldBboxMin = mpp::TPose2D(minx, miny, -M_PI);

    std::cout << "Start state: " << pi.stateStart.asString() << "\n";
    std::cout << "Goal state : " << pi.stateGoal.asString() << "\n";
    std::cout << "Obstacles  : " << obsPts->size() << " points\n";
    std::cout << "World bbox : (" << pi.worldBboxMin.x << "," << pi.worldBboxMin.y << ") - (" << pi.worldBboxMax.x << ","
              << pi.worldBboxMax.y << ")\n";

    const std::string plannerName = a.get("--planner", "mpp::TPS_Astar");

    mppb_ctx* ctx = nullptr;
    if (mppb_ctx_create(std::stoi(a.get("--device", "0")), nullptr, &ctx) != MPPB_OK)
        throw std::runtime_error(std::string("libmpp_b200: ") + mppb_last_global_error());
    struct CtxGuard
    {
        mppb_ctx* c;
        ~CtxGuard() { mppb_ctx_destroy(c); }
    };
    CtxGuard guard{ctx};  // destroyed after the planner's last use below (planner does not own ctx)
    // the planner class by name, from the class registry (the library's own classes and those of --plugins)
    std::shared_ptr<mpp::Planner> planner = mpp::ClassRegistry::instance().createPlanner(plannerName, ctx);
    if (!planner)
        throw std::runtime_error("Given classname '" + plannerName +
                                 "' does not seem to be a known C++ class implementing `Planner");

    if (a.isSet("--costmap-obstacles"))
    {
        const auto params = mpp::CostEvaluatorCostMap::Parameters::FromYAML(mpp::Yaml::FromFile(a.get("--costmap-obstacles")));
        planner->costEvaluators_.push_back(
            mpp::CostEvaluatorCostMap::FromStaticPointObstacles(ctx, *obsPts, params, pi.stateStart.pose));
    }
    auto wpParams = mpp::CostEvaluatorPreferredWaypoint::Parameters();
    if (a.isSet("--waypoints-parameters"))
        wpParams = mpp::CostEvaluatorPreferredWaypoint::Parameters::FromYAML(mpp::Yaml::FromFile(a.get("--waypoints-parameters")));
    if (a.isSet("--waypoints"))
    {
        auto ev     = std::make_shared<mpp::CostEvaluatorPreferredWaypoint>(ctx);
        ev->params_ = wpParams;
        ev->setPreferredWaypoints(mpp::waypoint_targets_from_yaml(mpp::Yaml::FromFile(a.get("--waypoints"))));
        planner->costEvaluators_.push_back(ev);
    }
    if (a.isSet("--planner-parameters"))
    {
        planner->params_from_yaml(mpp::Yaml::FromFile(a.get("--planner-parameters")));
        std::cout << "Loaded these planner params:\n" << planner->params_as_yaml_text();
    }
    planner->progressCallback_ = [](const mpp::ProgressCallbackData& pcd) {
        std::cout << "[progressCallback] bestCostFromStart: " << pcd.bestCostFromStart
                  << " bestCostToGoal: " << pcd.bestCostToGoal << " bestPathLength: " << pcd.bestPath.size() << std::endl;
    };

    std::cout << "[PTGs] Initializing PTGs..." << std::endl;
    pi.ptgs.initFromConfigFile(mpp::ConfigFile(a.get("--ptg-config")), a.get("--config-section", "SelfDriving"), ctx);
    std::cout << "[PTGs] Done." << std::endl;

    const mpp::PlannerOutput plan = planner->plan(pi);

    std::cout << "\nDone.\n";
    std::cout << "Success: " << (plan.success ? "YES" : "NO") << "\n";
    std::cout << "Plan has " << plan.motionTree.edges_to_children.size() << " overall edges, "
              << plan.motionTree.nodes().size() << " nodes\n";
    std::cout << "Plan time: " << plan.computationTime * 1e3 << " ms (" << plan.nodesExpanded << " nodes expanded, "
              << plan.tpsPointsEvaluated << " TPS points, " << plan.edgesCosted << " edges costed";
    if (auto* astar = dynamic_cast<mpp::TPS_Astar*>(planner.get()))
    {
        const auto& gs = astar->lastGpuStats;
        std::cout << "; GPU: " << gs.collisionBatches << "+" << gs.costBatches << " batches, " << gs.gpuSeconds * 1e3 << " ms";
    }
    std::cout << ")\n";
    if (plan.success) std::cout << "Path cost: " << plan.pathCost << "\n";

    if (!plan.bestNodeId.has_value())
    {
        std::cerr << "No bestNodeId in plan output.\n";
        return;
    }
    auto [plannedPath, pathEdges] = plan.motionTree.backtrack_path(*plan.bestNodeId);
    if (!a.isSet("--no-refine")) mpp::refine_trajectory(plannedPath, pathEdges, pi.ptgs);

    if (plan.success)
    {
        if (a.isSet("--print-path-edges"))
        {
            std::cout << "Planned path edges:\n";
            for (const auto& edge : pathEdges) std::cout << edge->asString();
        }
        if (a.isSet("--save-interpolated-path") || a.isSet("--play-animation"))
        {
            const double period = std::stod(a.get("--interpolarion-period", "0.25"));
            auto         traj   = mpp::plan_to_trajectory(pathEdges, pi.ptgs, period);
            const auto&  start  = plan.originalInput.stateStart.pose;
            for (auto& kv : traj) kv.second.state.pose = start + kv.second.state.pose;
            std::cout << "Interpolated path done.\n";
            if (a.isSet("--save-interpolated-path"))
            {
                std::cout << "Saving path to " << a.get("--save-interpolated-path") << std::endl;
                mpp::save_to_txt(traj, a.get("--save-interpolated-path"));
            }
        }
    }
    planner.reset();
}
}  // namespace

int main(int argc, char** argv)
{
    try
    {
        const Args a = parse(argc, argv);
        if (a.isSet("--help") || a.isSet("-h"))
        {
            usage();
            return 0;
        }
        if (a.isSet("--write-planner-parameters"))
        {
            std::ofstream f(a.get("--write-planner-parameters"));
            if (!f.is_open()) throw std::runtime_error("Cannot write " + a.get("--write-planner-parameters"));
            f << mpp::TPS_Astar_Parameters().as_yaml_text();
            std::cout << "Wrote file: " << a.get("--write-planner-parameters") << std::endl;
            return 0;
        }
        for (const char* req : {"--obstacles", "--ptg-config", "--start-pose", "--goal-pose"})
            if (!a.isSet(req))
            {
                std::cerr << "PARSE ERROR: Required argument missing: " << (req + 2) << "\n";
                usage();
                return 1;
            }
        if (a.isSet("--plugins"))
        {
            // shared objects whose initialisers register Planner / PTG classes (path-planner-cli.cpp:452-461)
            std::string err;
            if (!mpp::loadPluginModules(a.get("--plugins"), &err)) throw std::runtime_error("Could not load plugins, error: " + err);
        }
        do_plan_path(a);
        return 0;
    }
    catch (const std::exception& e)
    {
        std::cerr << e.what() << std::endl;
        return 1;
    }
}
