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
#include <iostream>
#include <map>
#include <memory>
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

#endif
