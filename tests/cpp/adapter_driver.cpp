/*
 * adapter_driver.cpp — TEST: the reference's own sources driving the GPU evaluator through the C++ adapter.
 *
 * Reads like the apps' mains ([REF] pendulum/src/Learn/main.cpp:14-45, stickgame/src/Learn/main.cpp:25-45): builds the
 * app's Instructions::Set with the reference's fillInstructionSet, instantiates the reference's LearningEnvironment
 * class, loads parameters, creates the agent, imports a golden graph ([REF] pendulum/src/CodeGen/mainTPGInference.cpp:28-34)
 * and calls evaluateAllRoots — with tpgx::LearningAgent (include/tpgx_gegelati.hpp) in place of
 * Learn::ParallelLearningAgent.  Prints one hex-float score per root; tests/test_gpu_cpp_adapter.py compares them with
 * the oracle.  Built by oracle/build_ref.sh (it compiles reference sources, so the binary lives under oracle/_ref/).
 */
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include <gegelati.h>

#include "pendulum/src/Learn/pendulum.h"
#include "gridworld/src/gridworld.h"
#include "stickgame/src/Learn/stickGameAdversarial.h"
#include "tic-tac-toe/src/Learn/TicTacToe.h"
#include "mnist/src/mnist.h"
#include "tpgx_gegelati.hpp"

void pendulum_fillInstructionSet(Instructions::Set& set);
void gridworld_fillInstructionSet(Instructions::Set& set);
void stickgame_fillInstructionSet(Instructions::Set& set);
static void tictactoe_fillInstructionSet(Instructions::Set& set) {
#include "gen/tictactoe_set.inc"
}
static void mnist_fillInstructionSet(Instructions::Set& set) {   /* [REF] mnist/src/main.cpp:47-88, sliced at build time */
#include "gen/mnist_set.inc"
}
/* the reference keeps its dataset protected: a subclass hands it to the adapter */
struct MNISTOpen : MNIST {
    static const mnist::MNIST_dataset<std::vector, std::vector<double>, uint8_t>& data() { return dataset; }
};

static int run_mnist(const char* dot, uint64_t generation, Learn::LearningMode mode, Learn::LearningParameters params) {
    Instructions::Set set;
    mnist_fillInstructionSet(set);
    MNISTOpen le;                                      /* [REF] mnist/src/main.cpp:101 */
    params.nbProgramConstant = 5;                      /* [REF] mnist/params.json:100 */
    const auto& ds = MNISTOpen::data();
    tpgx::ClassificationAgent la(le, set, params, ds.training_images, ds.training_labels, ds.test_images, ds.test_labels);
    Environment env(set, le.getDataSources(), params.nbRegisters, params.nbProgramConstant);
    File::TPGGraphDotImporter importer(dot, env, *la.getTPGGraph());
    importer.importGraph();
    auto results = la.evaluateAllRoots(generation, mode);
    for (const TPG::TPGVertex* r : la.getTPGGraph()->getRootVertices())
        for (const auto& kv : results)
            if (kv.second == r) {
                auto cr = std::dynamic_pointer_cast<Learn::ClassificationEvaluationResult>(kv.first);
                printf("%a %zu", cr->getResult(), cr->getNbEvaluation());
                for (double f : cr->getScorePerClass()) printf(" %a", f);
                printf("\n");
            }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s app dot generation mode iterations max_actions [second_player_random]\n", argv[0]); return 2; }
    const std::string app = argv[1];
    const uint64_t generation = strtoull(argv[3], nullptr, 10);
    const Learn::LearningMode mode = (Learn::LearningMode)atoi(argv[4]);
    Learn::LearningParameters params;
    params.nbRegisters = 8;
    params.nbIterationsPerPolicyEvaluation = strtoull(argv[5], nullptr, 10);
    params.maxNbActionsPerEval = strtoull(argv[6], nullptr, 10);
    const bool secondRandom = argc > 7 && atoi(argv[7]) != 0;
    try {
        if (app == "mnist") return run_mnist(argv[2], generation, mode, params);
        Instructions::Set set;
        std::unique_ptr<Learn::LearningEnvironment> le;
        tpgx_env_kind kind;
        const std::vector<double> actions = {0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.0};   /* [REF] pendulum/src/Learn/main.cpp:33 */
        if (app == "pendulum") { pendulum_fillInstructionSet(set); le.reset(new Pendulum(actions)); kind = TPGX_ENV_PENDULUM; params.nbProgramConstant = 5; }
        else if (app == "gridworld") { gridworld_fillInstructionSet(set); le.reset(new GridWorld()); kind = TPGX_ENV_GRIDWORLD; params.nbProgramConstant = 5; }
        else if (app == "stickgame") { stickgame_fillInstructionSet(set); le.reset(new StickGameAdversarial(secondRandom)); kind = TPGX_ENV_STICKGAME; }
        else if (app == "tictactoe") { tictactoe_fillInstructionSet(set); le.reset(new TicTacToe(secondRandom)); kind = TPGX_ENV_TICTACTOE; }
        else { fprintf(stderr, "unknown app\n"); return 2; }

        tpgx::LearningAgent la(*le, set, params, kind);
        la.setSecondPlayerRandom(secondRandom);
        if (app == "pendulum") la.setPendulumActions(actions);
        Environment env(set, le->getDataSources(), params.nbRegisters, params.nbProgramConstant);
        File::TPGGraphDotImporter dot(argv[2], env, *la.getTPGGraph());
        dot.importGraph();

        auto results = la.evaluateAllRoots(generation, mode);
        const auto& roots = la.getTPGGraph()->getRootVertices();
        for (const TPG::TPGVertex* r : roots)
            for (const auto& kv : results)
                if (kv.second == r) printf("%a %zu\n", kv.first->getResult(), kv.first->getNbEvaluation());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
