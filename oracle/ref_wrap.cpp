/*
 * oracle/ref_wrap.cpp — TEST INFRASTRUCTURE.  C wrapper around the reference's OWN sources.
 *
 * build_ref.sh compiles this file together with the unmodified reference translation units
 *   pendulum/src/Learn/{pendulum,instructions}.cpp      gridworld/src/{gridworld,instructions}.cpp
 *   stickgame/src/Learn/{stickGameAdversarial,instructions}.cpp   tic-tac-toe/src/Learn/TicTacToe.cpp
 * (taken where they lie under /root/reference, against oracle/shim/gegelati.h) and with the two
 * instruction-set blocks that live inside a main(): mnist/src/main.cpp and
 * tic-tac-toe/src/Learn/main.cpp, which the script slices at build time into oracle/_ref/gen/*.inc
 * (generated, git-ignored — no reference source is stored in this repository).
 * Output: oracle/_ref/libtpgref.so.  Used only by tests/ and tests/golden/make_fixtures.py to pin
 * the oracle's restatement of environment dynamics and instruction lambdas.
 */
#include <cfloat>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <numeric>

#include <gegelati.h>

#include "pendulum/src/Learn/pendulum.h"
#include "gridworld/src/gridworld.h"
#include "stickgame/src/Learn/stickGameAdversarial.h"
#include "tic-tac-toe/src/Learn/TicTacToe.h"

/* renamed per app by -D on the respective translation unit (build_ref.sh) */
void pendulum_fillInstructionSet(Instructions::Set& set);
void gridworld_fillInstructionSet(Instructions::Set& set);
void stickgame_fillInstructionSet(Instructions::Set& set);

static void mnist_fillInstructionSet(Instructions::Set& set) {
#include "gen/mnist_set.inc"
}
static void tictactoe_fillInstructionSet(Instructions::Set& set) {
#include "gen/tictactoe_set.inc"
}

namespace {

/* protected setters of the reference class, reachable from a subclass: lets a test start the
 * pendulum from an explicit state (SURVEY.md Appendix C.3 KAT) */
struct PendulumOpen : Pendulum {
    using Pendulum::Pendulum;
    void set(double angle, double velocity) { setAngle(angle); setVelocity(velocity); }
};

struct RefEnv {
    int kind;
    std::unique_ptr<Learn::LearningEnvironment> le;
};

Instructions::Set* get_set(int app) {
    static Instructions::Set sets[5];
    static bool done = false;
    if (!done) {
        mnist_fillInstructionSet(sets[0]);
        pendulum_fillInstructionSet(sets[1]);
        gridworld_fillInstructionSet(sets[2]);
        stickgame_fillInstructionSet(sets[3]);
        tictactoe_fillInstructionSet(sets[4]);
        done = true;
    }
    return (app >= 0 && app < 5) ? &sets[app] : nullptr;
}

} // namespace

extern "C" {

/* app: 0 mnist, 1 pendulum, 2 gridworld, 3 stickgame, 4 tic-tac-toe */
int ref_set_size(int app) { Instructions::Set* s = get_set(app); return s ? (int)s->getNbInstructions() : -1; }
int ref_set_nb_operands(int app, int i) { return (int)get_set(app)->getInstruction(i).getNbOperands(); }
/* 0 double, 1 int, 2 Data::Constant, 3 const double[3][3] */
int ref_set_operand_kind(int app, int i, int k) { return (int)get_set(app)->getInstruction(i).getOperandKind(k); }
/* d,i,c: one value per operand slot in each representation; win: 3x3 row-major */
double ref_set_call(int app, int i, const double* d, const int* iv, const double* c, const double* win) {
    double w[3][3];
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) w[r][k] = win ? win[r * 3 + k] : 0.0;
    return get_set(app)->getInstruction(i).call(d, iv, c, w);
}

/* kind: tpgx_env_kind (0 gridworld, 1 stickgame, 2 tic-tac-toe, 3 pendulum) */
void* ref_env_create(int kind, int second_player_random, int nb_pendulum_actions, const double* pendulum_actions) {
    RefEnv* e = new RefEnv{kind, nullptr};
    switch (kind) {
    case 0: e->le.reset(new GridWorld()); break;
    case 1: e->le.reset(new StickGameAdversarial(second_player_random != 0)); break;
    case 2: e->le.reset(new TicTacToe(second_player_random != 0)); break;
    case 3: e->le.reset(new PendulumOpen(std::vector<double>(pendulum_actions, pendulum_actions + nb_pendulum_actions))); break;
    default: delete e; return nullptr;
    }
    return e;
}
void ref_env_destroy(void* h) { delete static_cast<RefEnv*>(h); }
void* ref_env_clone(void* h) {
    RefEnv* e = static_cast<RefEnv*>(h);
    if (!e->le->isCopyable()) return nullptr;
    return new RefEnv{e->kind, std::unique_ptr<Learn::LearningEnvironment>(e->le->clone())};
}
void ref_env_reset(void* h, uint64_t seed, int mode) {
    static_cast<RefEnv*>(h)->le->reset((size_t)seed, (Learn::LearningMode)mode, 0, 0);
}
int ref_env_do_action(void* h, double action) {
    try { static_cast<RefEnv*>(h)->le->doAction(action); } catch (const std::exception&) { return -1; }
    return 0;
}
int ref_env_is_terminal(void* h) { return static_cast<RefEnv*>(h)->le->isTerminal() ? 1 : 0; }
double ref_env_score(void* h, int player) {
    RefEnv* e = static_cast<RefEnv*>(h);
    if (auto* adv = dynamic_cast<Learn::AdversarialLearningEnvironment*>(e->le.get())) return adv->getScores()->getScoreOf((size_t)player);
    return e->le->getScore();
}
int ref_env_nb_actions(void* h) { return (int)static_cast<RefEnv*>(h)->le->getNbActions(); }
/* what the programs see: every data source of getDataSources(), in order, every address, as double */
int ref_env_observe(void* h, double* out, int n) {
    RefEnv* e = static_cast<RefEnv*>(h);
    int k = 0;
    for (const Data::DataHandler& dh : e->le->getDataSources()) {
        const size_t nd = dh.getAddressSpace(typeid(double)), ni = dh.getAddressSpace(typeid(int));
        for (size_t a = 0; a < nd; a++, k++) if (k < n) out[k] = *dh.getDataAt(typeid(double), a).getSharedPointer<const double>();
        for (size_t a = 0; a < ni; a++, k++) if (k < n) out[k] = (double)*dh.getDataAt(typeid(int), a).getSharedPointer<const int>();
    }
    return k;
}
int ref_env_set_pendulum(void* h, double angle, double velocity) {
    RefEnv* e = static_cast<RefEnv*>(h);
    if (e->kind != 3) return -1;
    static_cast<PendulumOpen*>(e->le.get())->set(angle, velocity);
    return 0;
}

} /* extern "C" */
