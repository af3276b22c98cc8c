/*
 * oracle/shim/gegelati.h — TEST INFRASTRUCTURE, not product code.
 *
 * A from-scratch, minimal, API-compatible stand-in for the part of the GEGELATI library surface
 * (github.com/gegelati/gegelati, v2.0.0 per the golden dot headers) that the *in-tree* sources of
 * /root/reference touch: the five apps' LearningEnvironment subclasses and instruction-set
 * builders.  Its only purpose is to let oracle/build_ref.sh compile those
 * reference files UNMODIFIED, where they lie, into oracle/_ref/libtpgref.so, so that the CPU
 * oracle's restatement of the environment dynamics and of every instruction lambda can be pinned
 * against the reference's own arithmetic (tests/test_oracle_vs_ref.py, tests/golden/).
 *
 * Nothing here is copied from GEGELATI (the library is not available in this container); class and
 * method names follow the apps' call sites, bodies are the obvious minimal ones.  Behaviours that
 * are upstream assumptions are marked [U#]:
 *   [U9]  Data::Hash<T> = std::hash<T>
 *   [U10] Mutator::RNG  = std::mt19937_64 + std::uniform_{int,real}_distribution
 */
#ifndef TPGX_ORACLE_SHIM_GEGELATI_H
#define TPGX_ORACLE_SHIM_GEGELATI_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

namespace Data {

/* program constant operand: apps only convert it with (double)c ([REF] mnist/src/main.cpp:54) */
class Constant {
    double value;
  public:
    Constant(double v = 0.0) : value(v) {}
    operator double() const { return value; }
};

/* type-erased pointer returned by DataHandler::getDataAt; apps call getSharedPointer<const T>()
 * ([REF] pendulum/src/Learn/pendulum.cpp:27).  Scalars alias the handler's live storage (App. E ii). */
class UntypedSharedPtr {
    std::shared_ptr<const void> ptr;
  public:
    UntypedSharedPtr() = default;
    template <class T> explicit UntypedSharedPtr(std::shared_ptr<T> p) : ptr(std::move(p)) {}
    template <class T> std::shared_ptr<T> getSharedPointer() const {
        return std::shared_ptr<T>(ptr, static_cast<T*>(const_cast<void*>(ptr.get())));
    }
};

template <class T> struct Hash { /* [U9] */
    size_t operator()(const T& v) const { return std::hash<T>()(v); }
};

class DataHandler {
  public:
    virtual ~DataHandler() = default;
    virtual size_t getAddressSpace(const std::type_info& type) const = 0;
    virtual size_t getLargestAddressSpace() const = 0;
    virtual UntypedSharedPtr getDataAt(const std::type_info& type, size_t address) const = 0;
};

template <class T> class PrimitiveTypeArray : public DataHandler {
  protected:
    std::vector<T> data;
  public:
    explicit PrimitiveTypeArray(size_t n = 8) : data(n) {}
    PrimitiveTypeArray(const PrimitiveTypeArray&) = default; /* deep copy (App. E iii) */
    size_t getAddressSpace(const std::type_info& type) const override { return type == typeid(T) ? data.size() : 0; }
    size_t getLargestAddressSpace() const override { return data.size(); }
    UntypedSharedPtr getDataAt(const std::type_info& type, size_t address) const override {
        if (type != typeid(T) || address >= data.size()) throw std::invalid_argument("PrimitiveTypeArray::getDataAt");
        /* non-owning alias into live, contiguous storage */
        return UntypedSharedPtr(std::shared_ptr<const T>(std::shared_ptr<const T>(), &data[address]));
    }
    void setDataAt(const std::type_info& type, size_t address, const T& value) {
        if (type != typeid(T) || address >= data.size()) throw std::invalid_argument("PrimitiveTypeArray::setDataAt");
        data[address] = value;
    }
    void resetData() { for (T& v : data) v = T(); }
    const T* raw() const { return data.data(); } /* shim-only accessor for the wrapper */
};

/* 2-D view on a caller-owned vector ([REF] mnist/src/mnist.h:25, mnist.cpp:27): provides T (h*w addresses) and
 * T[3][3] windows ((h-2)*(w-2) addresses, address -> row a / (w-2), col a % (w-2)) */
template <class T> class Array2DWrapper : public DataHandler {
    size_t height, width;
    const std::vector<T>* pointer = nullptr;
  public:
    Array2DWrapper(size_t h, size_t w) : height(h), width(w) {}
    void setPointer(const std::vector<T>* p) { pointer = p; }
    size_t getAddressSpace(const std::type_info& type) const override {
        if (type == typeid(T)) return height * width;
        if (type == typeid(const T[3][3]) || type == typeid(T[3][3])) return (height - 2) * (width - 2);
        return 0;
    }
    size_t getLargestAddressSpace() const override { return height * width; }
    UntypedSharedPtr getDataAt(const std::type_info& type, size_t address) const override {
        if (!pointer) throw std::runtime_error("Array2DWrapper: null pointer");
        if (type == typeid(T)) return UntypedSharedPtr(std::shared_ptr<const T>(std::shared_ptr<const T>(), &pointer->at(address)));
        if (getAddressSpace(type) == 0 || address >= getAddressSpace(type)) throw std::invalid_argument("Array2DWrapper::getDataAt");
        const size_t r0 = address / (width - 2), c0 = address % (width - 2);
        std::shared_ptr<T> w(new T[9], std::default_delete<T[]>());
        for (size_t r = 0; r < 3; r++) for (size_t c = 0; c < 3; c++) w.get()[r * 3 + c] = pointer->at((r0 + r) * width + c0 + c);
        return UntypedSharedPtr(std::shared_ptr<const T>(w, w.get()));
    }
};

} // namespace Data

namespace Mutator {

class RNG { /* [U10] */
    std::mt19937_64 engine;
  public:
    RNG(uint64_t seed = 0) : engine(seed) {}
    void setSeed(uint64_t seed) { engine.seed(seed); }
    uint64_t getUnsignedInt64(uint64_t min, uint64_t max) { return std::uniform_int_distribution<uint64_t>(min, max)(engine); }
    int32_t getInt32(int32_t min, int32_t max) { return std::uniform_int_distribution<int32_t>(min, max)(engine); }
    double getDouble(double min, double max) { return std::uniform_real_distribution<double>(min, max)(engine); }
};

} // namespace Mutator

namespace Instructions {

/* operand type tags the wrapper can query */
enum class OperandKind { Double, Int, Constant, Window3x3 };

class Instruction {
  public:
    virtual ~Instruction() = default;
    virtual unsigned int getNbOperands() const = 0;
    virtual OperandKind getOperandKind(unsigned int i) const = 0;
    /* shim-only uniform call: every operand slot supplied in every representation */
    virtual double call(const double* d, const int* i, const double* c, const double (*w)[3]) const = 0;
    /* upstream-style API used by adapters: operand types and execution on type-erased arguments */
    virtual const std::vector<std::reference_wrapper<const std::type_info>>& getOperandTypes() const = 0;
    double execute(const std::vector<Data::UntypedSharedPtr>& args) const {
        double d[2] = {0, 0}, c[2] = {0, 0}, w[3][3] = {{0}};
        int iv[2] = {0, 0};
        for (unsigned k = 0; k < getNbOperands() && k < args.size(); k++) {
            switch (getOperandKind(k)) {
            case OperandKind::Double: d[k] = *args[k].getSharedPointer<const double>(); break;
            case OperandKind::Int: iv[k] = *args[k].getSharedPointer<const int>(); break;
            case OperandKind::Constant: c[k] = (double)*args[k].getSharedPointer<const Data::Constant>(); break;
            case OperandKind::Window3x3: { const double* p = args[k].getSharedPointer<const double>().get(); for (int r = 0; r < 9; r++) w[r / 3][r % 3] = p[r]; break; }
            }
        }
        return call(d, iv, c, w);
    }
};

namespace detail {
template <class T> struct kind_of;
template <> struct kind_of<double> { static constexpr OperandKind value = OperandKind::Double; };
template <> struct kind_of<int> { static constexpr OperandKind value = OperandKind::Int; };
template <> struct kind_of<Data::Constant> { static constexpr OperandKind value = OperandKind::Constant; };
template <> struct kind_of<const double[3][3]> { static constexpr OperandKind value = OperandKind::Window3x3; };

template <class T> struct arg_type { using type = T; };
template <> struct arg_type<const double[3][3]> { using type = const double (*)[3]; };

template <class T> struct fetch;
template <> struct fetch<double> { static double get(unsigned k, const double* d, const int*, const double*, const double (*)[3]) { return d[k]; } };
template <> struct fetch<int> { static int get(unsigned k, const double*, const int* i, const double*, const double (*)[3]) { return i[k]; } };
template <> struct fetch<Data::Constant> { static Data::Constant get(unsigned k, const double*, const int*, const double* c, const double (*)[3]) { return Data::Constant(c[k]); } };
template <> struct fetch<const double[3][3]> { static const double (*get(unsigned, const double*, const int*, const double*, const double (*w)[3]))[3] { return w; } };
} // namespace detail

template <class First, class... Rest> class LambdaInstruction : public Instruction {
    std::function<double(typename detail::arg_type<First>::type, typename detail::arg_type<Rest>::type...)> func;
    std::string printTemplate;
    template <size_t... Is>
    double invoke(std::index_sequence<Is...>, const double* d, const int* i, const double* c, const double (*w)[3]) const {
        return func(detail::fetch<First>::get(0, d, i, c, w), detail::fetch<Rest>::get((unsigned)(Is + 1), d, i, c, w)...);
    }
  public:
    template <class F> LambdaInstruction(F f, const std::string& tpl = "") : func(f), printTemplate(tpl) {}
    unsigned int getNbOperands() const override { return 1 + sizeof...(Rest); }
    OperandKind getOperandKind(unsigned int k) const override {
        const OperandKind kinds[] = {detail::kind_of<First>::value, detail::kind_of<Rest>::value...};
        return kinds[k];
    }
    double call(const double* d, const int* i, const double* c, const double (*w)[3]) const override {
        return invoke(std::index_sequence_for<Rest...>{}, d, i, c, w);
    }
    const std::vector<std::reference_wrapper<const std::type_info>>& getOperandTypes() const override {
        static const std::vector<std::reference_wrapper<const std::type_info>> types = {typeid(First), typeid(Rest)...};
        return types;
    }
    const std::string& getPrintTemplate() const { return printTemplate; }
};

class Set {
    std::vector<std::reference_wrapper<const Instruction>> instructions;
  public:
    bool add(const Instruction& ins) { instructions.emplace_back(ins); return true; }
    unsigned int getNbInstructions() const { return (unsigned int)instructions.size(); }
    const Instruction& getInstruction(uint64_t i) const { return instructions.at(i).get(); }
};

} // namespace Instructions

/* ---- minimal program / graph model (what an adapter flattens; names as the apps and upstream use them) ---- */
class Environment {
    const Instructions::Set& set;
    std::vector<std::reference_wrapper<const Data::DataHandler>> sources;
    size_t nbRegisters, nbConstants;
  public:
    Environment(const Instructions::Set& s, const std::vector<std::reference_wrapper<const Data::DataHandler>>& src, size_t nbRegs, size_t nbConst = 0)
        : set(s), sources(src), nbRegisters(nbRegs), nbConstants(nbConst) {}
    const Instructions::Set& getInstructionSet() const { return set; }
    const std::vector<std::reference_wrapper<const Data::DataHandler>>& getDataSources() const { return sources; }
    size_t getNbRegisters() const { return nbRegisters; }
    size_t getNbConstant() const { return nbConstants; }
};

namespace Program {
class Line {
    uint64_t instructionIndex = 0, destinationIndex = 0;
    std::pair<uint64_t, uint64_t> operands[2] = {{0, 0}, {0, 0}};
  public:
    uint64_t getInstructionIndex() const { return instructionIndex; }
    uint64_t getDestinationIndex() const { return destinationIndex; }
    const std::pair<uint64_t, uint64_t>& getOperand(unsigned k) const { return operands[k]; }
    void setInstructionIndex(uint64_t v) { instructionIndex = v; }
    void setDestinationIndex(uint64_t v) { destinationIndex = v; }
    void setOperand(unsigned k, uint64_t source, uint64_t location) { operands[k] = {source, location}; }
};
class Program {
    std::vector<Line> lines;
    std::vector<double> constants;
  public:
    explicit Program(const Environment& env) : constants(env.getNbConstant(), 0.0) {}
    size_t getNbLines() const { return lines.size(); }
    const Line& getLine(uint64_t i) const { return lines.at(i); }
    Line& addNewLine() { lines.emplace_back(); return lines.back(); }
    Data::Constant getConstantAt(size_t i) const { return Data::Constant(constants.at(i)); }
    void setConstantAt(size_t i, double v) { constants.at(i) = v; }
};
} // namespace Program

namespace TPG {
class TPGEdge;
class TPGVertex {
  protected:
    std::list<TPGEdge*> outgoing;
  public:
    virtual ~TPGVertex() = default;
    const std::list<TPGEdge*>& getOutgoingEdges() const { return outgoing; }
    void addOutgoingEdge(TPGEdge* e) { outgoing.push_back(e); }
};
class TPGTeam : public TPGVertex {};
class TPGAction : public TPGVertex {
    uint64_t actionID;
  public:
    explicit TPGAction(uint64_t id) : actionID(id) {}
    uint64_t getActionID() const { return actionID; }
};
class TPGEdge {
    const TPGVertex* source;
    const TPGVertex* destination;
    std::shared_ptr<Program::Program> program;
  public:
    TPGEdge(const TPGVertex* s, const TPGVertex* d, std::shared_ptr<Program::Program> p) : source(s), destination(d), program(std::move(p)) {}
    const TPGVertex* getSource() const { return source; }
    const TPGVertex* getDestination() const { return destination; }
    const Program::Program& getProgram() const { return *program; }
    std::shared_ptr<Program::Program> getProgramSharedPointer() const { return program; }
};
class TPGGraph {
    const Environment& env;
    std::vector<std::unique_ptr<TPGVertex>> vertices;
    std::list<std::unique_ptr<TPGEdge>> edges;
    std::vector<const TPGVertex*> roots;
  public:
    explicit TPGGraph(const Environment& e) : env(e) {}
    const Environment& getEnvironment() const { return env; }
    TPGTeam& addNewTeam() { vertices.emplace_back(new TPGTeam()); return static_cast<TPGTeam&>(*vertices.back()); }
    TPGAction& addNewAction(uint64_t id) { vertices.emplace_back(new TPGAction(id)); return static_cast<TPGAction&>(*vertices.back()); }
    TPGEdge& addNewEdge(TPGVertex& src, const TPGVertex& dst, std::shared_ptr<Program::Program> prog) {
        edges.emplace_back(new TPGEdge(&src, &dst, std::move(prog)));
        src.addOutgoingEdge(edges.back().get());
        return *edges.back();
    }
    std::vector<const TPGVertex*> getVertices() const { std::vector<const TPGVertex*> v; for (auto& x : vertices) v.push_back(x.get()); return v; }
    /* upstream derives roots from incoming edges; the importer records them from the file */
    void setRootVertices(const std::vector<const TPGVertex*>& r) { roots = r; }
    const std::vector<const TPGVertex*>& getRootVertices() const { return roots; }
    void clearProgramIntrons() {}
    void clear() { roots.clear(); edges.clear(); vertices.clear(); } /* [UP] TPGGraph::clear() */
};
} // namespace TPG
class Archive;
namespace TPG {
/* Declared so that code holding a CPU engine compiles ([REF] mnist/src/mnist.cpp:104-124); the shim has no CPU engine —
 * the repository's CPU restatement is oracle/tpg_oracle.cpp — so executing it is an error. */
class TPGExecutionEngine {
  public:
    TPGExecutionEngine(const Environment&, Archive* = nullptr) {}
    std::pair<std::vector<const TPGVertex*>, std::vector<size_t>> executeFromRoot(const TPGVertex&) {
        throw std::runtime_error("shim TPGExecutionEngine: no CPU engine behind this stand-in");
    }
};
} // namespace TPG

namespace File {
/* Reader of the TPGGraphDotExporter layout, enough for the reference's golden graphs. */
class TPGGraphDotImporter {
    std::string path;
    const Environment& env;
    TPG::TPGGraph& graph;
  public:
    TPGGraphDotImporter(const char* filePath, const Environment& e, TPG::TPGGraph& g) : path(filePath), env(e), graph(g) {}
    void importGraph() {
        std::ifstream in(path);
        if (!in) throw std::runtime_error("TPGGraphDotImporter: cannot open " + path);
        std::map<long, TPG::TPGTeam*> teams;
        std::map<long, std::shared_ptr<Program::Program>> programs;
        std::map<long, const TPG::TPGVertex*> progDest;
        std::map<long, long> actionLabel;
        std::string ln;
        auto num = [](const std::string& s, size_t& pos) { size_t used = 0; long v = std::stol(s.substr(pos), &used); pos += used; return v; };
        while (std::getline(in, ln)) {
            size_t p = ln.find_first_not_of(" \t");
            if (p == std::string::npos) continue;
            ln = ln.substr(p);
            if (ln.rfind("{ rank= same", 0) == 0) {
                std::vector<const TPG::TPGVertex*> roots;
                for (size_t q = ln.find('T'); q != std::string::npos; q = ln.find('T', q)) { q++; roots.push_back(teams.at(num(ln, q))); }
                graph.setRootVertices(roots);
            } else if (ln[0] == 'T' && ln.find("[fillcolor") != std::string::npos) {
                size_t q = 1; teams[num(ln, q)] = &graph.addNewTeam();
            } else if (ln[0] == 'P' && ln.find("shape=point") != std::string::npos) {
                size_t q = 1; const long id = num(ln, q);
                auto prog = std::make_shared<Program::Program>(env);
                q = ln.find("//");
                if (q != std::string::npos) {
                    q += 2; size_t i = 0;
                    while (q < ln.size() && i < env.getNbConstant()) { size_t bar = ln.find('|', q); if (bar == std::string::npos) break; prog->setConstantAt(i++, std::stod(ln.substr(q, bar - q))); q = bar + 1; }
                }
                programs[id] = prog;
            } else if (ln[0] == 'I' && ln.find("label=\"") != std::string::npos) {
                size_t q = 1; const long id = num(ln, q);
                q = ln.find("label=\"") + 7;
                const size_t end = ln.find('"', q);
                std::string body = ln.substr(q, end - q);
                size_t b = 0;
                while (b < body.size()) {
                    size_t e2 = body.find("&#92;n", b);
                    if (e2 == std::string::npos) e2 = body.size();
                    std::string item = body.substr(b, e2 - b);
                    if (!item.empty()) {   /* instr|dest&src0|loc0#src1|loc1 */
                        long v[6]; size_t r = 0;
                        for (int k = 0; k < 6; k++) { v[k] = num(item, r); r++; }
                        Program::Line& L = programs.at(id)->addNewLine();
                        L.setInstructionIndex((uint64_t)v[0]); L.setDestinationIndex((uint64_t)v[1]);
                        L.setOperand(0, (uint64_t)v[2], (uint64_t)v[3]); L.setOperand(1, (uint64_t)v[4], (uint64_t)v[5]);
                    }
                    b = e2 + 6;
                }
            } else if (ln[0] == 'A' && ln.find("label=\"") != std::string::npos) {
                size_t q = 1; const long id = num(ln, q);
                q = ln.find("label=\"") + 7; actionLabel[id] = num(ln, q);
            } else if (ln[0] == 'T' && ln.find("-> P") != std::string::npos) {
                size_t q = 1; const long t = num(ln, q);
                q = ln.find("-> P") + 4; const long pid = num(ln, q);
                size_t arrow = ln.find("->", q);
                if (arrow != std::string::npos) {
                    q = ln.find_first_of("TA", arrow);
                    const char kind = ln[q]; q++;
                    const long d = num(ln, q);
                    progDest[pid] = kind == 'T' ? static_cast<const TPG::TPGVertex*>(teams.at(d)) : &graph.addNewAction((uint64_t)actionLabel.at(d));
                }
                graph.addNewEdge(*teams.at(t), *progDest.at(pid), programs.at(pid));
            }
        }
    }
};
} // namespace File

namespace Learn {

enum class LearningMode { TRAINING, VALIDATION, TESTING };

class LearningEnvironment {
  protected:
    const uint64_t nbActions;
  public:
    LearningEnvironment() = delete;
    LearningEnvironment(uint64_t nbAct) : nbActions(nbAct) {}
    LearningEnvironment(const LearningEnvironment&) = default;
    virtual ~LearningEnvironment() = default;
    virtual LearningEnvironment* clone() const { return nullptr; }
    virtual bool isCopyable() const { return false; }
    uint64_t getNbActions() const { return nbActions; }
    /* the upstream base implementation only validates the id */
    virtual void doAction(double actionID) {
        if (actionID < 0 || actionID >= (double)nbActions) throw std::runtime_error("Given action ID exceeds the number of actions for this environment.");
    }
    virtual void reset(size_t seed = 0, LearningMode mode = LearningMode::TRAINING, uint16_t iterationNumber = 0, uint64_t generationNumber = 0) = 0;
    virtual std::vector<std::reference_wrapper<const Data::DataHandler>> getDataSources() = 0;
    virtual double getScore() const = 0;
    virtual bool isTerminal() const = 0;
};

class EvaluationResult {
  protected:
    double result;
    size_t nbEvaluation;
  public:
    EvaluationResult(double res = 0.0, size_t nbEval = 1) : result(res), nbEvaluation(nbEval) {}
    virtual ~EvaluationResult() = default;
    virtual double getResult() const { return result; }
    virtual size_t getNbEvaluation() const { return nbEvaluation; }
};

class ClassificationLearningEnvironment : public LearningEnvironment {   /* [UP]; App. F U11 */
  protected:
    std::vector<std::vector<uint64_t>> classificationTable;
    uint64_t currentClass = 0;
  public:
    ClassificationLearningEnvironment(uint64_t nbClass) : LearningEnvironment(nbClass), classificationTable(nbClass, std::vector<uint64_t>(nbClass, 0)) {}
    void doAction(double actionID) override {
        LearningEnvironment::doAction(actionID);
        classificationTable.at(currentClass).at((uint64_t)actionID)++;
    }
    const std::vector<std::vector<uint64_t>>& getClassificationTable() const { return classificationTable; }
    void reset(size_t = 0, LearningMode = LearningMode::TRAINING, uint16_t = 0, uint64_t = 0) override {
        for (auto& row : classificationTable) std::fill(row.begin(), row.end(), 0);
    }
    double getScore() const override {
        double sum = 0.0;
        const size_t C = classificationTable.size();
        for (size_t c = 0; c < C; c++) {
            uint64_t tp = classificationTable[c][c], fn = 0, fp = 0;
            for (size_t k = 0; k < C; k++) { if (k != c) { fn += classificationTable[c][k]; fp += classificationTable[k][c]; } }
            const double recall = (double)tp / (double)(tp + fn), precision = (double)tp / (double)(tp + fp);
            sum += (tp != 0) ? 2 * (precision * recall) / (precision + recall) : 0.0;
        }
        return sum / (double)C;
    }
};

class ClassificationEvaluationResult : public EvaluationResult {
    std::vector<double> scorePerClass;
    std::vector<size_t> nbEvaluationPerClass;
  public:
    ClassificationEvaluationResult(const std::vector<double>& scores, const std::vector<size_t>& nbEval)
        : EvaluationResult(scores.empty() ? 0.0 : std::accumulate(scores.begin(), scores.end(), 0.0) / (double)scores.size(),
                           std::accumulate(nbEval.begin(), nbEval.end(), (size_t)0)),
          scorePerClass(scores), nbEvaluationPerClass(nbEval) {}
    const std::vector<double>& getScorePerClass() const { return scorePerClass; }
    const std::vector<size_t>& getNbEvaluationPerClass() const { return nbEvaluationPerClass; }
};

struct TPGParameters {   /* [UP] Mutator::TPGParameters, the keys of "mutation": {"tpg": {...}} */
    size_t nbRoots = 100, maxInitOutgoingEdges = 3, maxOutgoingEdges = 5;
    double pEdgeDeletion = 0.7, pEdgeAddition = 0.7, pProgramMutation = 0.2, pEdgeDestinationChange = 0.1, pEdgeDestinationIsAction = 0.5;
    bool forceProgramBehaviorChangeOnMutation = false;
};
struct ProgramParameters {   /* [UP] Mutator::ProgramParameters, "mutation": {"prog": {...}} */
    size_t maxProgramSize = 96;
    double pDelete = 0.5, pAdd = 0.5, pMutate = 1.0, pSwap = 1.0, pConstantMutation = 0.5;
    int32_t minConstValue = -100, maxConstValue = 100;
};
struct MutationParameters { TPGParameters tpg; ProgramParameters prog; };
struct LearningParameters {   /* the fields the apps' params.json files set */
    MutationParameters mutation;
    size_t nbRegisters = 8, nbProgramConstant = 0, archiveSize = 50;
    double archivingProbability = 0.05, ratioDeletedRoots = 0.5;
    uint64_t maxNbActionsPerEval = 1000, nbIterationsPerPolicyEvaluation = 5, nbIterationsPerJob = 1, maxNbEvaluationPerPolicy = 1000;
    uint64_t nbGenerations = 1, nbThreads = 1;
    bool doValidation = false;
};

class AdversarialEvaluationResult : public EvaluationResult {
    std::vector<double> scores;
  public:
    AdversarialEvaluationResult(const std::vector<double>& s, size_t nbEval = 1) : EvaluationResult(s.empty() ? 0.0 : s[0], nbEval), scores(s) {}
    double getScoreOf(size_t i) const { return scores.at(i); }
    size_t getSize() const { return scores.size(); }
};

class AdversarialLearningEnvironment : public LearningEnvironment {
  public:
    AdversarialLearningEnvironment(uint64_t nbAct) : LearningEnvironment(nbAct) {}
    virtual std::shared_ptr<AdversarialEvaluationResult> getScores() const = 0;
    double getScore() const override { return getScores()->getScoreOf(0); }
};

} // namespace Learn

/* ---- scaffolding for compiling an app's main.cpp UNMODIFIED ([REF] gridworld/src/main.cpp): parameter file reader, graph
 * exporter, loggers and policy statistics are upstream subsystems that are out of scope; what follows is the
 * least that lets the unmodified main() compile and run, with Learn::ParallelLearningAgent standing for the GPU-backed agent
 * of include/tpgx_gegelati.hpp (in a real GEGELATI build the maintainer subclasses / patches ParallelLearningAgent instead:
 * INTEGRATION.md).  Enabled by -DTPGX_SHIM_WITH_AGENT. */
#ifdef TPGX_SHIM_WITH_AGENT
#include "tpgx_gegelati.hpp"
namespace Learn {
using LearningAgent = tpgx::TrainingAgent;
using ParallelLearningAgent = tpgx::TrainingAgent;
} // namespace Learn
namespace File {
class ParametersParser {
    static std::string strip(const std::string& in) { /* the apps' params.json files carry C and C++ comments */
        std::string out;
        for (size_t i = 0; i < in.size();) {
            if (in.compare(i, 2, "/*") == 0) { const size_t e = in.find("*/", i + 2); i = e == std::string::npos ? in.size() : e + 2; }
            else if (in.compare(i, 2, "//") == 0) { const size_t e = in.find('\n', i); i = e == std::string::npos ? in.size() : e; }
            else out += in[i++];
        }
        return out;
    }
    static bool find(const std::string& txt, const std::string& key, std::string* val) {
        const size_t k = txt.find("\"" + key + "\"");
        if (k == std::string::npos) return false;
        size_t c = txt.find(':', k);
        if (c == std::string::npos) return false;
        c++;
        while (c < txt.size() && isspace((unsigned char)txt[c])) c++;
        size_t e = c;
        while (e < txt.size() && txt[e] != ',' && txt[e] != '}' && txt[e] != '\n') e++;
        *val = txt.substr(c, e - c);
        return true;
    }
    template <class T> static void num(const std::string& txt, const char* key, T& field) {
        std::string v;
        if (!find(txt, key, &v)) return;
        if (v.find("true") != std::string::npos) field = (T)1;
        else if (v.find("false") != std::string::npos) field = (T)0;
        else field = (T)std::stod(v);
    }
  public:
    static void loadParametersFromJson(const char* path, Learn::LearningParameters& p) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error(std::string("cannot open ") + path);
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string t = strip(ss.str());
        num(t, "archiveSize", p.archiveSize); num(t, "archivingProbability", p.archivingProbability); num(t, "doValidation", p.doValidation);
        num(t, "maxNbActionsPerEval", p.maxNbActionsPerEval); num(t, "maxNbEvaluationPerPolicy", p.maxNbEvaluationPerPolicy);
        num(t, "nbGenerations", p.nbGenerations); num(t, "nbIterationsPerPolicyEvaluation", p.nbIterationsPerPolicyEvaluation);
        num(t, "nbIterationsPerJob", p.nbIterationsPerJob); num(t, "nbProgramConstant", p.nbProgramConstant); num(t, "nbRegisters", p.nbRegisters);
        num(t, "nbThreads", p.nbThreads); num(t, "ratioDeletedRoots", p.ratioDeletedRoots);
        auto& g = p.mutation.tpg;
        num(t, "nbRoots", g.nbRoots); num(t, "maxInitOutgoingEdges", g.maxInitOutgoingEdges); num(t, "maxOutgoingEdges", g.maxOutgoingEdges);
        num(t, "pEdgeDeletion", g.pEdgeDeletion); num(t, "pEdgeAddition", g.pEdgeAddition); num(t, "pProgramMutation", g.pProgramMutation);
        num(t, "pEdgeDestinationChange", g.pEdgeDestinationChange); num(t, "pEdgeDestinationIsAction", g.pEdgeDestinationIsAction);
        num(t, "forceProgramBehaviorChangeOnMutation", g.forceProgramBehaviorChangeOnMutation);
        auto& m = p.mutation.prog;
        num(t, "maxProgramSize", m.maxProgramSize); num(t, "pDelete", m.pDelete); num(t, "pAdd", m.pAdd); num(t, "pMutate", m.pMutate); num(t, "pSwap", m.pSwap);
        num(t, "pConstantMutation", m.pConstantMutation); num(t, "minConstValue", m.minConstValue); num(t, "maxConstValue", m.maxConstValue);
    }
    static void writeParametersToJson(const char*, const Learn::LearningParameters&) {}
};
class TPGGraphDotExporter { /* writes vertex / edge counts only: the product's exporter is python/tpgx/graph.py export_dot */
    std::string path;
    const TPG::TPGGraph& graph;
  public:
    TPGGraphDotExporter(const char* p, const TPG::TPGGraph& g) : path(p), graph(g) {}
    void setNewFilePath(const char* p) { path = p; }
    void print() {
        std::ofstream f(path);
        f << "// stand-in exporter: " << graph.getVertices().size() << " vertices, " << graph.getRootVertices().size() << " roots\n";
    }
};
} // namespace File
namespace Log {
class LABasicLogger { public: explicit LABasicLogger(Learn::LearningAgent& la, std::ostream& os = std::cout) { la.logTo(&os); } };
class LAPolicyStatsLogger { public: LAPolicyStatsLogger(Learn::LearningAgent&, std::ostream&) {} };
} // namespace Log
namespace TPG {
class PolicyStats {
    const TPGVertex* root = nullptr;
  public:
    void setEnvironment(const Environment&) {}
    void analyzePolicy(const TPGVertex* r) { root = r; }
    friend std::ostream& operator<<(std::ostream& os, const PolicyStats& ps) { return os << "policy root " << (ps.root ? "set" : "missing") << "\n"; }
};
} // namespace TPG
#endif

#endif
