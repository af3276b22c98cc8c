/*
 * tpgx_gegelati.hpp — the reference-side binding: a header-only C++17 adapter between GEGELATI's own
 * types (<gegelati.h>) and the C ABI of libtpgx.so (tpgx.h).
 *
 * tpgx::LearningAgent mirrors the part of Learn::LearningAgent / ParallelLearningAgent the apps use
 * around evaluation ([REF] pendulum/src/Learn/main.cpp:38-39,83; tic-tac-toe/src/Learn/main.cpp:82-83):
 * same constructor arguments, getTPGGraph(), and
 *     evaluateAllRoots(generationNumber, mode) -> std::multimap<std::shared_ptr<Learn::EvaluationResult>,
 *                                                                 const TPG::TPGVertex*>
 * with the hot path running on the GPU.  It needs only public GEGELATI API: Instructions::Set /
 * Instruction::{getNbOperands,getOperandTypes,execute}, TPG::TPGGraph::{getVertices,getRootVertices},
 * TPGVertex::getOutgoingEdges, TPGEdge::{getProgram,getDestination}, TPGAction::getActionID,
 * Program::{getNbLines,getLine,getConstantAt}, Line::{getInstructionIndex,getDestinationIndex,getOperand}.
 * GEGELATI itself is not vendored in /root/reference; in this repository the header is compiled against
 * the API-compatible stand-in oracle/shim/gegelati.h (tests/test_gpu_cpp_adapter.py).
 *
 * Errors: any non-zero tpgx_status is rethrown as std::runtime_error(tpgx_last_error()), matching the
 * reference's error behaviour ([REF] pendulum/src/Learn/pendulum.cpp:66-68).  No CPU fallback: an
 * instruction whose behaviour matches no device opcode, or an environment the device does not
 * implement, is an error.
 */
#ifndef TPGX_GEGELATI_HPP
#define TPGX_GEGELATI_HPP

#include <gegelati.h>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "tpgx.h"

namespace tpgx {

inline void check(tpgx_ctx* ctx, tpgx_status st, const char* where) {
    if (st != TPGX_OK) throw std::runtime_error(std::string(where) + ": " + tpgx_last_error(ctx));
}

/* ---- Instructions::Set -> device opcodes by behavioural fingerprint ------------------------------
 * LambdaInstruction wraps an opaque lambda, so each instruction is run on a fixed probe set and its
 * results are compared, bit for bit, with the host evaluation of every device behaviour that has the
 * same operand signature.  Transcendentals are compared through std:: (what the
 * lambdas call); the device computes them with its own single-source libm (csrc/tpgx_math.h).      */
namespace detail {
inline double host_behaviour(int op, double a, double b, int ia, int ib, double c, const double* w) {
    auto gx = [&] { return -w[0] + w[2] - 2.0 * w[3] + 2.0 * w[5] - w[6] + w[8]; };
    auto gy = [&] { return -w[0] - 2.0 * w[1] - w[2] + w[6] + 2.0 * w[7] + w[8]; };
    switch (op) {
    case TPGX_OP_SUB: return a - b;
    case TPGX_OP_ADD: return a + b;
    case TPGX_OP_MUL: return a * b;
    case TPGX_OP_DIV: return a / b;
    case TPGX_OP_MAX: return (a < b) ? b : a;
    case TPGX_OP_EXP: return std::exp(a);
    case TPGX_OP_LOG: return std::log(a);
    case TPGX_OP_COS: return std::cos(a);
    case TPGX_OP_SIN: return std::sin(a);
    case TPGX_OP_TAN: return std::tan(a);
    case TPGX_OP_PI: return M_PI;
    case TPGX_OP_MULC: return a * c / 10.0;
    case TPGX_OP_FMODG: return b != 0.0 ? std::fmod(a, b) : DBL_MIN;
    case TPGX_OP_SUB_II: return (double)ia - (double)ib;
    case TPGX_OP_CAST_I: return (double)ia;
    case TPGX_OP_EQ0: return a == 0.0 ? 10.0 : 0.0;
    case TPGX_OP_EQM1: return a == -1.0 ? 10.0 : 0.0;
    case TPGX_OP_EQ1: return a == 1.0 ? 10.0 : 0.0;
    case TPGX_OP_GE15: return a >= 15.0 ? 10.0 : 0.0;
    case TPGX_OP_COND: return a < b ? -a : a;
    case TPGX_OP_SOBEL_MAGN: { const double x = gx(), y = gy(); return std::sqrt(x * x + y * y); }
    case TPGX_OP_SOBEL_DIR: return std::atan(gy() / gx());
    default: return 0.0;
    }
}
inline const char* signature_of(int op) { /* D double, I int, C Data::Constant, W const double[3][3] */
    switch (op) {
    case TPGX_OP_SUB: case TPGX_OP_ADD: case TPGX_OP_MUL: case TPGX_OP_DIV: case TPGX_OP_MAX: case TPGX_OP_FMODG: case TPGX_OP_COND: return "DD";
    case TPGX_OP_MULC: return "DC";
    case TPGX_OP_SUB_II: return "II";
    case TPGX_OP_CAST_I: return "I";
    case TPGX_OP_SOBEL_MAGN: case TPGX_OP_SOBEL_DIR: return "W";
    default: return "D";
    }
}
inline bool same_bits(double x, double y) { return (x != x && y != y) || std::memcmp(&x, &y, 8) == 0; }
} // namespace detail

inline std::vector<int32_t> fingerprint(const Instructions::Set& set) {
    static const double PD[][2] = {{0.0, 0.0}, {1.0, 2.0}, {-1.0, 1.0}, {0.5, -3.25}, {15.0, 7.0}, {16.5, -0.0}, {-7.75, 2.5}, {1e-3, 1e3},
                                   {3.0, 0.0}, {0.0, 5.0}, {2.75, 2.75}, {-1.0, -1.0}, {100.25, -0.125}, {1.0, 1.0}, {0.3, 0.7}, {20.0, 3.0}};
    static const int PI_[][2] = {{0, 0}, {3, 1}, {-4, 9}, {21, 4}, {7, 7}, {1, -1}, {100, 3}, {-8, -8}, {5, 2}, {0, 6}, {9, 0}, {2, 2}, {13, 5}, {1, 1}, {4, 8}, {6, 3}};
    static const double PW[][9] = {{0, 0, 0, 0, 0, 0, 0, 0, 0}, {1, 2, 3, 4, 5, 6, 7, 8, 9}, {255, 0, 0, 0, 255, 0, 0, 0, 255}, {10, 0, 200, 30, 7, 90, 0, 0, 1},
                                   {0, 255, 0, 255, 0, 255, 0, 255, 0}, {9, 8, 7, 6, 5, 4, 3, 2, 1}, {0, 0, 255, 0, 0, 255, 0, 0, 255}, {3, 1, 4, 1, 5, 9, 2, 6, 5}};
    std::vector<int32_t> ops;
    for (uint64_t i = 0; i < set.getNbInstructions(); i++) {
        const Instructions::Instruction& ins = set.getInstruction(i);
        std::string sig;
        for (const std::type_info& t : ins.getOperandTypes())
            sig += t == typeid(double) ? 'D' : t == typeid(int) ? 'I' : t == typeid(Data::Constant) ? 'C' : t == typeid(const double[3][3]) ? 'W' : '?';
        int found = -1;
        for (int op = 0; op < TPGX_OP_COUNT && found < 0; op++) {
            if (sig != detail::signature_of(op)) continue;
            bool all = true;
            const int n = sig == "W" ? 8 : 16;
            for (int k = 0; k < n && all; k++) {
                std::vector<Data::UntypedSharedPtr> args;
                const double d0 = PD[k % 16][0], d1 = PD[k % 16][1];
                for (size_t s = 0; s < sig.size(); s++) {
                    if (sig[s] == 'D') args.emplace_back(std::make_shared<const double>(s == 0 ? d0 : d1));
                    else if (sig[s] == 'I') args.emplace_back(std::make_shared<const int>(PI_[k][s]));
                    else if (sig[s] == 'C') args.emplace_back(std::make_shared<const Data::Constant>(Data::Constant(d1)));
                    else { /* 3x3 window: 9 contiguous doubles */
                        std::shared_ptr<double> w(new double[9], std::default_delete<double[]>());
                        std::memcpy(w.get(), PW[k % 8], sizeof(double) * 9);
                        args.emplace_back(std::shared_ptr<const double>(w, w.get()));
                    }
                }
                const double got = ins.execute(args);
                const double want = detail::host_behaviour(op, d0, d1, PI_[k][0], PI_[k][1], d1, PW[k % 8]);
                all = detail::same_bits(got, want);
            }
            if (all) found = op;
        }
        if (found < 0) throw std::runtime_error("tpgx::fingerprint: instruction " + std::to_string(i) + " matches no device opcode (no CPU fallback)");
        ops.push_back(found);
    }
    return ops;
}

/* ---- TPG::TPGGraph -> tpgx_graph, in the library's own iteration order -------------------------- */
struct FlatGraph {
    std::vector<int64_t> lineOffset;
    std::vector<tpgx_line> lines;
    std::vector<double> constants;
    std::vector<int32_t> teamEdgeOffset, edgeProgram, edgeDestination, rootTeam;
    std::vector<const TPG::TPGVertex*> roots;
    int32_t nbPrograms = 0, nbTeams = 0;
    tpgx_graph view() const {
        return tpgx_graph{nbPrograms, lineOffset.data(), lines.data(), nullptr, constants.empty() ? nullptr : constants.data(), nbTeams,
                          teamEdgeOffset.data(), edgeProgram.data(), edgeDestination.data(), (int32_t)rootTeam.size(), rootTeam.data()};
    }
};

inline FlatGraph flatten(const TPG::TPGGraph& graph) {
    FlatGraph f;
    const size_t nc = graph.getEnvironment().getNbConstant();
    std::map<const TPG::TPGVertex*, int32_t> teamId;
    std::map<const Program::Program*, int32_t> progId;
    std::vector<const TPG::TPGTeam*> teams;
    for (const TPG::TPGVertex* v : graph.getVertices())
        if (auto* t = dynamic_cast<const TPG::TPGTeam*>(v)) { teamId[v] = (int32_t)teams.size(); teams.push_back(t); }
    f.lineOffset.push_back(0);
    f.teamEdgeOffset.push_back(0);
    for (const TPG::TPGTeam* t : teams) {
        for (const TPG::TPGEdge* e : t->getOutgoingEdges()) {          /* std::list order: it decides ties */
            const Program::Program* p = &e->getProgram();
            auto it = progId.find(p);
            if (it == progId.end()) {
                it = progId.emplace(p, (int32_t)progId.size()).first;
                for (uint64_t l = 0; l < p->getNbLines(); l++) {
                    const Program::Line& L = p->getLine(l);
                    tpgx_line out{(uint32_t)L.getInstructionIndex(), (uint32_t)L.getDestinationIndex(),
                                  {(uint32_t)L.getOperand(0).first, (uint32_t)L.getOperand(1).first},
                                  {(uint32_t)L.getOperand(0).second, (uint32_t)L.getOperand(1).second}};
                    f.lines.push_back(out);
                }
                f.lineOffset.push_back((int64_t)f.lines.size());
                for (size_t c = 0; c < nc; c++) f.constants.push_back((double)p->getConstantAt(c));
            }
            f.edgeProgram.push_back(it->second);
            const TPG::TPGVertex* d = e->getDestination();
            if (auto* a = dynamic_cast<const TPG::TPGAction*>(d)) f.edgeDestination.push_back(-1 - (int32_t)a->getActionID());
            else f.edgeDestination.push_back(teamId.at(d));
        }
        f.teamEdgeOffset.push_back((int32_t)f.edgeProgram.size());
    }
    for (const TPG::TPGVertex* r : graph.getRootVertices())
        if (teamId.count(r)) { f.rootTeam.push_back(teamId.at(r)); f.roots.push_back(r); }
    f.nbPrograms = (int32_t)progId.size();
    f.nbTeams = (int32_t)teams.size();
    return f;
}

/* ---- the agent ----------------------------------------------------------------------------------- */
class LearningAgent {
  protected:
    Learn::LearningEnvironment& learningEnvironment;
    Learn::LearningParameters params;
    Environment env;
    std::shared_ptr<TPG::TPGGraph> tpg;
    tpgx_ctx* ctx = nullptr;
    tpgx_env_kind kind;
    int secondPlayerRandom = 0;
    std::vector<double> pendulumActions;

  public:
    /* `kind` names the device restatement of `le` (the app's LearningEnvironment subclass) */
    LearningAgent(Learn::LearningEnvironment& le, const Instructions::Set& set, const Learn::LearningParameters& p, tpgx_env_kind k,
                  int tieBreak = TPGX_TIE_LAST_MAX)
        : learningEnvironment(le), params(p), env(set, le.getDataSources(), p.nbRegisters, p.nbProgramConstant),
          tpg(std::make_shared<TPG::TPGGraph>(env)), kind(k) {
        tpgx_config cfg{};
        cfg.device = 0; cfg.nb_registers = (int32_t)p.nbRegisters; cfg.nb_constants = (int32_t)p.nbProgramConstant; cfg.tie_break = tieBreak;
        check(nullptr, tpgx_create(&cfg, &ctx), "tpgx_create");
        const std::vector<int32_t> ops = fingerprint(set);
        check(ctx, tpgx_set_instruction_table(ctx, (int32_t)ops.size(), ops.data()), "tpgx_set_instruction_table");
        std::vector<tpgx_source_desc> src;
        for (const Data::DataHandler& dh : le.getDataSources()) {   /* PrimitiveTypeArray<double|int>(n) */
            const size_t nd = dh.getAddressSpace(typeid(double)), ni = dh.getAddressSpace(typeid(int));
            src.push_back(tpgx_source_desc{nd ? TPGX_ELEM_F64 : TPGX_ELEM_I32, 1, (int32_t)(nd ? nd : ni), 0});
        }
        check(ctx, tpgx_set_data_sources(ctx, (int32_t)src.size(), src.data()), "tpgx_set_data_sources");
    }
    LearningAgent(const LearningAgent&) = delete;
    ~LearningAgent() { tpgx_destroy(ctx); }

    std::shared_ptr<TPG::TPGGraph> getTPGGraph() { return tpg; }
    void setSecondPlayerRandom(bool r) { secondPlayerRandom = r ? 1 : 0; }
    void setPendulumActions(const std::vector<double>& a) { pendulumActions = a; }   /* [REF] pendulum/src/Learn/main.cpp:33 */

    /* [UP] LearningAgent::evaluateAllRoots: one job per root, nbIterationsPerPolicyEvaluation episodes of at most
       maxNbActionsPerEval actions each, episode seed Hash(generationNumber) ^ Hash(iteration), score = mean. */
    std::multimap<std::shared_ptr<Learn::EvaluationResult>, const TPG::TPGVertex*> evaluateAllRoots(uint64_t generationNumber, Learn::LearningMode mode) {
        const FlatGraph f = flatten(*tpg);
        const tpgx_graph g = f.view();
        check(ctx, tpgx_set_graph(ctx, &g), "tpgx_set_graph");
        const int NI = (int)params.nbIterationsPerPolicyEvaluation, R = (int)f.roots.size();
        std::vector<uint64_t> seeds(NI), rngSeeds(NI);
        for (int it = 0; it < NI; it++) {
            seeds[it] = Data::Hash<uint64_t>()(generationNumber) ^ Data::Hash<uint64_t>()((uint64_t)it);
            /* what the environment's reset() hands to rng.setSeed() ([REF] pendulum/src/Learn/pendulum.cpp:45-48), computed here with the
               linked library's Data::Hash so that libtpgx does not have to know it */
            rngSeeds[it] = (uint64_t)(Data::Hash<size_t>()((size_t)seeds[it]) ^ Data::Hash<Learn::LearningMode>()(mode));
        }
        std::vector<int32_t> jobRoots(R);
        for (int r = 0; r < R; r++) jobRoots[r] = r;
        tpgx_episode_request rq{};
        rq.env = kind; rq.mode = (int32_t)mode; rq.nb_iterations = NI; rq.max_actions = (int32_t)params.maxNbActionsPerEval;
        rq.seeds = seeds.data(); rq.rng_seeds = rngSeeds.data(); rq.nb_jobs = R; rq.roots_per_job = 1; rq.job_roots = jobRoots.data();
        rq.second_player_random = secondPlayerRandom;
        rq.nb_pendulum_actions = (int32_t)pendulumActions.size(); rq.pendulum_actions = pendulumActions.data();
        std::vector<double> score(R);
        tpgx_episode_result out{};
        out.score = score.data();
        check(ctx, tpgx_evaluate_episodes(ctx, &rq, &out), "tpgx_evaluate_episodes");
        std::multimap<std::shared_ptr<Learn::EvaluationResult>, const TPG::TPGVertex*> results;
        for (int r = 0; r < R; r++) results.emplace(std::make_shared<Learn::EvaluationResult>(score[r], (size_t)NI), f.roots[r]);
        return results;
    }
};

/* Classification (MNIST-shaped) agent: mirrors Learn::ClassificationLearningAgent<ParallelLearningAgent>
 * ([REF] mnist/src/main.cpp:106).  The environment keeps its dataset private ([REF] mnist/src/mnist.h:16), so the two image
 * sets are handed over once (converted to uint8: pixels are the integers 0..255 stored as double).  evaluateAllRoots restates
 * [UP] ClassificationLearningAgent::evaluateJob with nbIterationsPerPolicyEvaluation = 1 ([REF] mnist/params.json:97): the
 * images are those MNIST::reset / doAction would draw for seed Hash(generation) ^ Hash(0) (tpgx_mnist_episode_indices). */
class ClassificationAgent {
    Learn::LearningParameters params;
    Environment env;
    std::shared_ptr<TPG::TPGGraph> tpg;
    tpgx_ctx* ctx = nullptr;
    int64_t nbTrain = 0, nbTest = 0;
    int nbClasses;

  public:
    ClassificationAgent(Learn::ClassificationLearningEnvironment& le, const Instructions::Set& set, const Learn::LearningParameters& p,
                        const std::vector<std::vector<double>>& trainImages, const std::vector<uint8_t>& trainLabels,
                        const std::vector<std::vector<double>>& testImages, const std::vector<uint8_t>& testLabels, int tieBreak = TPGX_TIE_LAST_MAX)
        : params(p), env(set, le.getDataSources(), p.nbRegisters, p.nbProgramConstant), tpg(std::make_shared<TPG::TPGGraph>(env)),
          nbTrain((int64_t)trainImages.size()), nbTest((int64_t)testImages.size()), nbClasses((int)le.getNbActions()) {
        tpgx_config cfg{};
        cfg.device = 0; cfg.nb_registers = (int32_t)p.nbRegisters; cfg.nb_constants = (int32_t)p.nbProgramConstant; cfg.tie_break = tieBreak;
        check(nullptr, tpgx_create(&cfg, &ctx), "tpgx_create");
        const std::vector<int32_t> ops = fingerprint(set);
        check(ctx, tpgx_set_instruction_table(ctx, (int32_t)ops.size(), ops.data()), "tpgx_set_instruction_table");
        /* the 2-D source's shape from its public address spaces: h*w scalars, (h-2)*(w-2) 3x3 windows (square image) */
        const Data::DataHandler& dh = le.getDataSources().at(0).get();
        const size_t A_ = dh.getAddressSpace(typeid(double)), B_ = dh.getAddressSpace(typeid(const double[3][3]));
        const size_t sum = (A_ - B_ + 4) / 2;                       /* h + w */
        const size_t side = sum / 2;
        if (side * side != A_ || (side - 2) * (side - 2) != B_) throw std::runtime_error("tpgx::ClassificationAgent: a square 2-D double source is expected");
        tpgx_source_desc src{TPGX_ELEM_F64, (int32_t)side, (int32_t)side, 1};
        check(ctx, tpgx_set_data_sources(ctx, 1, &src), "tpgx_set_data_sources");
        std::vector<uint8_t> img((size_t)(nbTrain + nbTest) * A_), lab((size_t)(nbTrain + nbTest));
        auto put = [&](const std::vector<std::vector<double>>& images, const std::vector<uint8_t>& labels, size_t at) {
            for (size_t i = 0; i < images.size(); i++) {
                for (size_t k = 0; k < A_; k++) {
                    const double v = images[i].at(k);
                    if (!(v >= 0.0 && v <= 255.0) || v != std::floor(v)) throw std::runtime_error("tpgx::ClassificationAgent: pixel is not an integer in [0, 255]");
                    img[(at + i) * A_ + k] = (uint8_t)v;
                }
                lab[at + i] = labels.at(i);
            }
        };
        put(trainImages, trainLabels, 0);
        put(testImages, testLabels, (size_t)nbTrain);
        check(ctx, tpgx_set_observations(ctx, 0, TPGX_STORE_U8, img.data(), lab.data(), nbTrain + nbTest), "tpgx_set_observations");
    }
    ClassificationAgent(const ClassificationAgent&) = delete;
    ~ClassificationAgent() { tpgx_destroy(ctx); }
    std::shared_ptr<TPG::TPGGraph> getTPGGraph() { return tpg; }

    std::multimap<std::shared_ptr<Learn::EvaluationResult>, const TPG::TPGVertex*> evaluateAllRoots(uint64_t generationNumber, Learn::LearningMode mode) {
        const FlatGraph f = flatten(*tpg);
        const tpgx_graph g = f.view();
        check(ctx, tpgx_set_graph(ctx, &g), "tpgx_set_graph");
        const bool training = mode == Learn::LearningMode::TRAINING;
        const int64_t size = training ? nbTrain : nbTest, n = (int64_t)params.maxNbActionsPerEval;
        std::vector<int32_t> idx((size_t)n);
        const uint64_t seed = Data::Hash<uint64_t>()(generationNumber) ^ Data::Hash<uint64_t>()((uint64_t)0);
        if (tpgx_mnist_episode_indices(seed, (int32_t)mode, size, n, idx.data()) != TPGX_OK) throw std::runtime_error("tpgx_mnist_episode_indices");
        if (!training) for (int32_t& i : idx) i += (int32_t)nbTrain;       /* the test set follows the training set on the device */
        const int R = (int)f.roots.size(), C = nbClasses;
        tpgx_classif_request rq{n, idx.data(), C, 0, R, 0};
        std::vector<double> f1((size_t)R * C);
        std::vector<uint64_t> conf((size_t)R * C * C);
        tpgx_classif_result out{};
        out.class_f1 = f1.data();
        out.confusion = conf.data();
        check(ctx, tpgx_evaluate_classification(ctx, &rq, &out), "tpgx_evaluate_classification");
        std::multimap<std::shared_ptr<Learn::EvaluationResult>, const TPG::TPGVertex*> results;
        for (int r = 0; r < R; r++) {
            std::vector<size_t> perClass(C, 0);
            for (int c = 0; c < C; c++) for (int k = 0; k < C; k++) perClass[c] += (size_t)conf[((size_t)r * C + c) * C + k];
            results.emplace(std::make_shared<Learn::ClassificationEvaluationResult>(std::vector<double>(f1.begin() + (size_t)r * C, f1.begin() + (size_t)(r + 1) * C), perClass),
                            f.roots[r]);
        }
        return results;
    }
};

/* ---- the training surface the apps call ------------------------------------------------------------
 * [REF] gridworld/src/main.cpp:34-36,63,68,79 / pendulum/src/Learn/main.cpp:38-39,83,88,96 (v2 spelling) and
 * stickgame/src/Learn/main.cpp:70,81 (old spelling): init(), trainOneGeneration(i), getTPGGraph(),
 * getSelector()->{getBestRoot, keepBestPolicy}, getBestRoot(), keepBestPolicy().  The generation loop itself is
 * libtpgx's host trainer (tpgx_trainer_*: populate / evaluate / archive / skip-and-merge / decimate / validation) with
 * every evaluation on the GPU; this class owns the trainer, feeds it the GPU backend and keeps the GEGELATI-side
 * TPG::TPGGraph object (the one exporters and loggers hold a reference to) in step with it.                         */
inline tpgx_env_kind envKindOf(const Learn::LearningEnvironment& le) {
    const std::string n = typeid(le).name(); /* the mangled name holds the class name */
    if (n.find("GridWorld") != std::string::npos) return TPGX_ENV_GRIDWORLD;
    if (n.find("Pendulum") != std::string::npos) return TPGX_ENV_PENDULUM;
    if (n.find("StickGame") != std::string::npos) return TPGX_ENV_STICKGAME;
    if (n.find("TicTacToe") != std::string::npos) return TPGX_ENV_TICTACTOE;
    throw std::runtime_error("tpgx: no device restatement of LearningEnvironment " + n + " (no CPU fallback)");
}

class TrainingAgent : public LearningAgent {
    tpgx_trainer* trainer = nullptr;
    tpgx_train_backend backend{};
    std::vector<tpgx_source_desc> sources;
    std::vector<int32_t> opcodes;
    int obsLen = 0;
    std::ostream* log = nullptr;
    tpgx_train_stats last{};
    std::vector<const TPG::TPGVertex*> rootVertex; /* GEGELATI-side vertex of every current root, in root order */

    /* backend callbacks: evaluateAllRoots(generation, mode) on the GPU with upstream's seed derivation */
    static tpgx_status evaluateCb(void* user, const tpgx_graph* g, uint64_t generation, int32_t mode, const uint64_t* archiveSeed, double probability,
                                  int32_t archiveSize, double* scores, int32_t, double*, uint64_t*, int32_t* recProgram, double* recBid, double* recObs,
                                  int32_t obsLen_, int32_t* nbRecords) {
        TrainingAgent* a = static_cast<TrainingAgent*>(user);
        tpgx_status st = tpgx_set_graph(a->ctx, g);
        if (st != TPGX_OK) return st;
        const int R = g->nb_roots, NI = (int)a->params.nbIterationsPerPolicyEvaluation;
        std::vector<int32_t> jobRoots((size_t)R);
        for (int j = 0; j < R; j++) jobRoots[j] = j;
        std::vector<uint64_t> seeds((size_t)NI), rngSeeds((size_t)NI);
        for (int it = 0; it < NI; it++) {
            seeds[it] = Data::Hash<uint64_t>()(generation) ^ Data::Hash<uint64_t>()((uint64_t)it);
            rngSeeds[it] = (uint64_t)(Data::Hash<size_t>()((size_t)seeds[it]) ^ Data::Hash<Learn::LearningMode>()((Learn::LearningMode)mode));
        }
        tpgx_episode_request rq{};
        rq.env = a->kind; rq.mode = mode; rq.nb_iterations = NI; rq.max_actions = (int32_t)a->params.maxNbActionsPerEval;
        rq.seeds = seeds.data(); rq.rng_seeds = rngSeeds.data(); rq.nb_jobs = R; rq.roots_per_job = 1; rq.job_roots = jobRoots.data();
        rq.second_player_random = a->secondPlayerRandom;
        rq.nb_pendulum_actions = (int32_t)a->pendulumActions.size(); rq.pendulum_actions = a->pendulumActions.data();
        tpgx_episode_result ev{};
        ev.score = scores;
        *nbRecords = 0;
        if (archiveSize <= 0 || probability == 0.0 || mode != 0) return tpgx_evaluate_episodes(a->ctx, &rq, &ev);
        const size_t cap = (size_t)archiveSize;
        std::vector<int32_t> job(cap), iteration(cap), step(cap);
        tpgx_archive_request ar{archiveSeed, probability, archiveSize, 0};
        tpgx_episode_archive_result out{recProgram, job.data(), iteration.data(), step.data(), recBid, recObs, obsLen_, 0};
        st = tpgx_archive_episodes(a->ctx, &rq, &ar, &ev, &out);
        *nbRecords = out.nb_records;
        return st;
    }
    static tpgx_status executeCb(void* user, const tpgx_graph* g, int64_t count, const double* obs, int32_t obsLen_, double* bid) {
        TrainingAgent* a = static_cast<TrainingAgent*>(user);
        tpgx_status st = tpgx_set_graph(a->ctx, g);
        if (st != TPGX_OK) return st;
        std::vector<std::vector<double>> f(a->sources.size());
        std::vector<std::vector<int32_t>> iv(a->sources.size());
        const double* pf[2] = {nullptr, nullptr};
        const int32_t* pi[2] = {nullptr, nullptr};
        int col = 0;
        for (size_t k = 0; k < a->sources.size(); k++) { /* archived observation rows -> one array per data source */
            const int sp = a->sources[k].height * a->sources[k].width;
            if (a->sources[k].elem == TPGX_ELEM_F64) {
                f[k].resize((size_t)count * sp);
                for (int64_t i = 0; i < count; i++) for (int q = 0; q < sp; q++) f[k][(size_t)i * sp + q] = obs[(size_t)i * obsLen_ + col + q];
                pf[k] = f[k].data();
            } else {
                iv[k].resize((size_t)count * sp);
                for (int64_t i = 0; i < count; i++) for (int q = 0; q < sp; q++) iv[k][(size_t)i * sp + q] = (int32_t)obs[(size_t)i * obsLen_ + col + q];
                pi[k] = iv[k].data();
            }
            col += sp;
        }
        return tpgx_execute_programs(a->ctx, count, pf, pi, bid);
    }
    void checkTrainer(tpgx_status st, const char* where) const {
        if (st != TPGX_OK) throw std::runtime_error(std::string(where) + ": " + (trainer ? tpgx_trainer_last_error(trainer) : "no trainer") + " / " + tpgx_last_error(ctx));
    }
    /* the trainer's graph -> the GEGELATI-side graph object, rebuilt in place (vertices in creation order: teams, then actions) */
    void syncGraph() {
        tpgx_graph v{};
        checkTrainer(tpgx_trainer_graph(trainer, &v), "tpgx_trainer_graph");
        tpg->clear();
        std::vector<TPG::TPGTeam*> team((size_t)v.nb_teams);
        for (int t = 0; t < v.nb_teams; t++) team[t] = &tpg->addNewTeam();
        std::map<int, const TPG::TPGAction*> action;
        std::vector<std::shared_ptr<Program::Program>> prog((size_t)v.nb_programs);
        const size_t nc = env.getNbConstant();
        for (int p = 0; p < v.nb_programs; p++) {
            prog[p] = std::make_shared<Program::Program>(env);
            for (int64_t l = v.program_line_offset[p]; l < v.program_line_offset[p + 1]; l++) {
                Program::Line& L = prog[p]->addNewLine();
                L.setInstructionIndex(v.lines[l].instruction);
                L.setDestinationIndex(v.lines[l].destination);
                L.setOperand(0, v.lines[l].src[0], v.lines[l].loc[0]);
                L.setOperand(1, v.lines[l].src[1], v.lines[l].loc[1]);
            }
            for (size_t c = 0; c < nc; c++) prog[p]->setConstantAt(c, v.constants[(size_t)p * nc + c]);
        }
        for (int t = 0; t < v.nb_teams; t++)
            for (int e = v.team_edge_offset[t]; e < v.team_edge_offset[t + 1]; e++) {
                const int d = v.edge_destination[e];
                const TPG::TPGVertex* dst;
                if (d >= 0) dst = team[d];
                else {
                    auto it = action.find(-1 - d);
                    if (it == action.end()) it = action.emplace(-1 - d, &tpg->addNewAction((uint64_t)(-1 - d))).first;
                    dst = it->second;
                }
                tpg->addNewEdge(*team[t], *dst, prog[v.edge_program[e]]);
            }
        rootVertex.clear();
        for (int r = 0; r < v.nb_roots; r++) rootVertex.push_back(team[v.root_team[r]]);
        tpg->setRootVertices(rootVertex);
    }

  public:
    TrainingAgent(Learn::LearningEnvironment& le, const Instructions::Set& set, const Learn::LearningParameters& p, int tieBreak = TPGX_TIE_LAST_MAX)
        : LearningAgent(le, set, p, envKindOf(le), tieBreak) {
        opcodes = fingerprint(set);
        for (const Data::DataHandler& dh : le.getDataSources()) {
            const size_t nd = dh.getAddressSpace(typeid(double)), ni = dh.getAddressSpace(typeid(int));
            sources.push_back(tpgx_source_desc{nd ? TPGX_ELEM_F64 : TPGX_ELEM_I32, 1, (int32_t)(nd ? nd : ni), 0});
            obsLen += (int)(nd ? nd : ni);
        }
        backend.user = this; backend.evaluate = &TrainingAgent::evaluateCb; backend.execute = &TrainingAgent::executeCb;
    }
    ~TrainingAgent() { if (trainer) tpgx_trainer_destroy(trainer); }

    /* [UP] LearningAgent::init(seed): RNG seeded, initRandomTPG */
    void init(uint64_t seed = 0) {
        if (trainer) { tpgx_trainer_destroy(trainer); trainer = nullptr; }
        tpgx_config cfg{};
        cfg.device = 0; cfg.nb_registers = (int32_t)params.nbRegisters; cfg.nb_constants = (int32_t)params.nbProgramConstant; cfg.tie_break = TPGX_TIE_LAST_MAX;
        const auto& mt = params.mutation.tpg;
        const auto& mp = params.mutation.prog;
        tpgx_train_params tp{};
        tp.nb_actions = (int32_t)learningEnvironment.getNbActions(); tp.nb_roots = (int32_t)mt.nbRoots;
        tp.max_init_outgoing_edges = (int32_t)mt.maxInitOutgoingEdges; tp.max_outgoing_edges = (int32_t)mt.maxOutgoingEdges;
        tp.p_edge_deletion = mt.pEdgeDeletion; tp.p_edge_addition = mt.pEdgeAddition; tp.p_program_mutation = mt.pProgramMutation;
        tp.p_edge_destination_change = mt.pEdgeDestinationChange; tp.p_edge_destination_is_action = mt.pEdgeDestinationIsAction;
        tp.force_program_behavior_change = mt.forceProgramBehaviorChangeOnMutation ? 1 : 0;
        tp.max_program_size = (int32_t)mp.maxProgramSize; tp.p_delete = mp.pDelete; tp.p_add = mp.pAdd; tp.p_mutate = mp.pMutate; tp.p_swap = mp.pSwap;
        tp.p_constant_mutation = mp.pConstantMutation; tp.min_const_value = (int32_t)mp.minConstValue; tp.max_const_value = (int32_t)mp.maxConstValue;
        tp.ratio_deleted_roots = params.ratioDeletedRoots; tp.archive_size = (int32_t)params.archiveSize; tp.archiving_probability = params.archivingProbability;
        tp.max_nb_evaluation_per_policy = (int32_t)params.maxNbEvaluationPerPolicy; tp.do_validation = params.doValidation ? 1 : 0;
        tp.nb_iterations_per_policy_evaluation = (int32_t)params.nbIterationsPerPolicyEvaluation; tp.nb_classes = 0;
        checkTrainer(tpgx_trainer_create(&cfg, (int32_t)opcodes.size(), opcodes.data(), (int32_t)sources.size(), sources.data(), &tp, seed, &trainer), "tpgx_trainer_create");
        syncGraph();
    }
    /* [UP] LearningAgent::trainOneGeneration */
    void trainOneGeneration(uint64_t generationNumber) {
        if (!trainer) throw std::runtime_error("tpgx::TrainingAgent::trainOneGeneration: init() has not been called");
        checkTrainer(tpgx_trainer_train_generation(trainer, &backend, generationNumber, &last), "tpgx_trainer_train_generation");
        syncGraph();
        if (log) {
            char buf[256];
            snprintf(buf, sizeof buf, "%6llu %6d %8d %12.4f %12.4f %12.4f %6d %6d %016llx\n", (unsigned long long)generationNumber, last.nb_teams, last.nb_programs,
                     last.worst_score, last.mean_score, last.best_score, last.nb_evaluated_roots, last.nb_skipped_roots, (unsigned long long)tpgx_trainer_fingerprint(trainer));
            *log << buf << std::flush;
        }
    }
    const tpgx_train_stats& getLastStats() const { return last; }
    uint64_t fingerprintOfGraph() const { return trainer ? tpgx_trainer_fingerprint(trainer) : 0; }
    void logTo(std::ostream* os) {
        log = os;
        if (log) *log << "   Gen  Teams Programs          Min          Avg          Max   Jobs Skipped      Fingerprint\n";
    }
    /* [UP] getBestRoot(): (vertex, its stored result) */
    std::pair<const TPG::TPGVertex*, std::shared_ptr<Learn::EvaluationResult>> getBestRoot() const {
        int32_t idx = -1;
        double score = 0.0;
        if (!trainer || tpgx_trainer_best_root(trainer, &idx, &score) != TPGX_OK || idx < 0 || idx >= (int32_t)rootVertex.size()) return {nullptr, nullptr};
        std::vector<double> sc(rootVertex.size());
        std::vector<uint64_t> ne(rootVertex.size());
        tpgx_trainer_root_results(trainer, sc.data(), ne.data(), (int32_t)sc.size());
        return {rootVertex[(size_t)idx], std::make_shared<Learn::EvaluationResult>(score, (size_t)ne[(size_t)idx])};
    }
    /* [UP] keepBestPolicy() */
    void keepBestPolicy() {
        checkTrainer(tpgx_trainer_keep_best_policy(trainer), "tpgx_trainer_keep_best_policy");
        syncGraph();
    }
    /* v2 spelling: la.getSelector()->getBestRoot() / keepBestPolicy() */
    TrainingAgent* getSelector() { return this; }
};

} // namespace tpgx
#endif
