// preset.cpp — Qt-free preset load/save: the host side of CellFlowWidget::loadPreset / savePreset
// (cuda-native/src/CellFlowWidget.cpp:1070-1269).  Same keys, same "only if present" rule, same
// double -> float narrowing (QJsonValue::toDouble assigned to float fields).  Unknown keys (e.g.
// "metaball" in settings.json) are ignored, as QJsonObject lookups by known key ignore them.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/cellflow_b200.h"

namespace {

struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const {
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool lit(const char* s) {
        size_t n = strlen(s);
        if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    std::string parse_string() {
        std::string out;
        if (p >= end || *p != '"') {
            ok = false;
            return out;
        }
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                        if (end - p >= 5) {
                            unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                            out += cp < 0x80 ? (char)cp : '?';
                            p += 4;
                        }
                        break;
                    default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) ok = false; else p++;
        return out;
    }
    JValue parse_value(int depth = 0) {
        JValue v;
        ws();
        if (p >= end || depth > 64) {
            ok = false;
            return v;
        }
        if (*p == '{') {
            v.kind = JValue::Object;
            p++;
            ws();
            if (p < end && *p == '}') {
                p++;
                return v;
            }
            while (ok) {
                ws();
                std::string key = parse_string();
                ws();
                if (p >= end || *p != ':') {
                    ok = false;
                    break;
                }
                p++;
                JValue child = parse_value(depth + 1);
                v.obj.emplace_back(std::move(key), std::move(child));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == '}') {
                    p++;
                    break;
                }
                ok = false;
            }
        } else if (*p == '[') {
            v.kind = JValue::Array;
            p++;
            ws();
            if (p < end && *p == ']') {
                p++;
                return v;
            }
            while (ok) {
                v.arr.push_back(parse_value(depth + 1));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == ']') {
                    p++;
                    break;
                }
                ok = false;
            }
        } else if (*p == '"') {
            v.kind = JValue::String;
            v.str = parse_string();
        } else if (lit("true")) {
            v.kind = JValue::Bool;
            v.b = true;
        } else if (lit("false")) {
            v.kind = JValue::Bool;
            v.b = false;
        } else if (lit("null")) {
            v.kind = JValue::Null;
        } else {
            char* e = nullptr;
            v.num = strtod(p, &e);
            if (e == p || e > end) {
                ok = false;
            } else {
                v.kind = JValue::Number;
                p = e;
            }
        }
        return v;
    }
};

// QJsonValue::toDouble(): the number, or 0 for any other kind; toInt(): the number if it is
// integral, else 0; toBool(): the bool, else false.
double to_double(const JValue* v) { return v && v->kind == JValue::Number ? v->num : 0.0; }
int to_int(const JValue* v) {
    if (!v || v->kind != JValue::Number) return 0;
    double d = v->num;
    return d == std::floor(d) && std::fabs(d) < 2147483648.0 ? (int)d : 0;
}
int to_bool(const JValue* v) { return v && v->kind == JValue::Bool && v->b ? 1 : 0; }

}  // namespace

extern "C" void cf_default_preset(cf_preset* pr) {
    if (!pr) return;
    memset(pr, 0, sizeof(*pr));
    cf_default_params(&pr->params);
    pr->particleCount = 4000; /* CellFlowWidget ctor, CellFlowWidget.cpp:13 */
    pr->pointSize = 10.0f;    /* SimulationParams.h:40-56 */
    pr->depthFadeStart = 10000.0f;
    pr->depthFadeEnd = 15000.0f;
    pr->sizeAttenuationFactor = 1000.0f;
    pr->brightnessMin = 0.4f;
    pr->focusDistance = 3000.0f;
    pr->apertureSize = 0.0f;
    pr->enableDepthFade = 0;
    pr->enableSizeAttenuation = 1;
    pr->enableBrightnessAttenuation = 1;
    pr->enableDOF = 0;
}

extern "C" int cf_load_preset(const char* path, cf_preset* pr) {
    if (!path || !pr) return CF_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return CF_ERR_IO; /* loadPreset returns false, CellFlowWidget.cpp:1072-1074 */
    std::string text;
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    fclose(f);
    Parser ps{text.data(), text.data() + text.size()};
    JValue root = ps.parse_value();
    if (!ps.ok || root.kind != JValue::Object) return CF_ERR_IO;

    cf_params& p = pr->params;
    const JValue* v;
#define F(key, field) if ((v = root.get(key))) field = (float)to_double(v)
#define B(key, field) if ((v = root.get(key))) field = to_bool(v)
    if ((v = root.get("PARTICLE_COUNT"))) pr->particleCount = to_int(v);
    if ((v = root.get("numParticleTypes"))) p.numParticleTypes = to_int(v);
    F("radius", p.radius);
    F("delta_t", p.delta_t);
    F("friction", p.friction);
    F("repulsion", p.repulsion);
    F("attraction", p.attraction);
    F("k", p.k);
    F("balance", p.balance);
    F("forceMultiplier", p.forceMultiplier);
    F("forceRange", p.forceRange);
    F("forceBias", p.forceBias);
    F("ratio", p.ratio);
    F("lfoA", p.lfoA);
    F("lfoS", p.lfoS);
    F("forceOffset", p.forceOffset);
    F("pointSize", pr->pointSize);
    F("canvasWidth", p.canvasWidth);
    F("canvasHeight", p.canvasHeight);
    F("canvasDepth", p.canvasDepth);
    F("spawnRegionSize", p.spawnRegionSize);
    F("depthFadeStart", pr->depthFadeStart);
    F("depthFadeEnd", pr->depthFadeEnd);
    F("sizeAttenuationFactor", pr->sizeAttenuationFactor);
    F("brightnessMin", pr->brightnessMin);
    F("focusDistance", pr->focusDistance);
    F("apertureSize", pr->apertureSize);
    B("enableDepthFade", pr->enableDepthFade);
    B("enableSizeAttenuation", pr->enableSizeAttenuation);
    B("enableBrightnessAttenuation", pr->enableBrightnessAttenuation);
    B("enableDOF", pr->enableDOF);
    B("invertPan", pr->invertPan);
    B("invertForwardBack", pr->invertForwardBack);
    B("invertRotation", pr->invertRotation);
    if ((v = root.get("effectType"))) pr->effectType = to_int(v);
#undef F
#undef B
    if (p.numParticleTypes < 1 || p.numParticleTypes > CF_MAX_PARTICLE_TYPES) return CF_ERR_IO;
    p.ratioWithLFO = p.ratio; /* lfoA == 0 path of paintGL, CellFlowWidget.cpp:419-421 */
    if ((v = root.get("particleColors")) && v->kind == JValue::Array) {
        int n = 0;
        for (size_t i = 0; i < v->arr.size() && i < CF_MAX_PARTICLE_TYPES; i++) {
            const JValue& c = v->arr[i];
            pr->particleColors[i].r = (float)to_double(c.get("r"));
            pr->particleColors[i].g = (float)to_double(c.get("g"));
            pr->particleColors[i].b = (float)to_double(c.get("b"));
            n++;
        }
        pr->numColors = n;
    }
    if ((v = root.get("radioByType")) && v->kind == JValue::Array) {
        int n = 0;
        for (size_t i = 0; i < v->arr.size() && (int)i < p.numParticleTypes; i++)
            pr->radioByType[n++] = (float)to_double(&v->arr[i]);
        pr->numRadio = n;
    }
    if ((v = root.get("rawForceTable")) && v->kind == JValue::Array) {
        int n = 0, lim = p.numParticleTypes * p.numParticleTypes;
        for (size_t i = 0; i < v->arr.size() && (int)i < lim; i++)
            pr->rawForceTable[n++] = (float)to_double(&v->arr[i]);
        pr->numRawForce = n;
    }
    return CF_OK;
}

// savePreset (CellFlowWidget.cpp:1182-1269): same key set; numbers written with enough digits
// (%.17g of the float widened to double) to round-trip exactly, as QJsonDocument does.
extern "C" int cf_save_preset(const char* path, const cf_preset* pr) {
    if (!path || !pr) return CF_ERR_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return CF_ERR_IO;
    const cf_params& p = pr->params;
    int T = p.numParticleTypes;
    fprintf(f, "{\n");
    fprintf(f, "    \"PARTICLE_COUNT\": %d,\n", pr->particleCount);
    fprintf(f, "    \"numParticleTypes\": %d,\n", T);
#define F(key, val) fprintf(f, "    \"%s\": %.17g,\n", key, (double)(val))
#define B(key, val) fprintf(f, "    \"%s\": %s,\n", key, (val) ? "true" : "false")
    F("radius", p.radius);
    F("delta_t", p.delta_t);
    F("friction", p.friction);
    F("repulsion", p.repulsion);
    F("attraction", p.attraction);
    F("k", p.k);
    F("balance", p.balance);
    F("forceMultiplier", p.forceMultiplier);
    F("forceRange", p.forceRange);
    F("forceBias", p.forceBias);
    F("ratio", p.ratio);
    F("lfoA", p.lfoA);
    F("lfoS", p.lfoS);
    F("forceOffset", p.forceOffset);
    F("pointSize", pr->pointSize);
    F("canvasWidth", p.canvasWidth);
    F("canvasHeight", p.canvasHeight);
    F("canvasDepth", p.canvasDepth);
    F("spawnRegionSize", p.spawnRegionSize);
    F("depthFadeStart", pr->depthFadeStart);
    F("depthFadeEnd", pr->depthFadeEnd);
    F("sizeAttenuationFactor", pr->sizeAttenuationFactor);
    F("brightnessMin", pr->brightnessMin);
    F("focusDistance", pr->focusDistance);
    F("apertureSize", pr->apertureSize);
    B("enableDepthFade", pr->enableDepthFade);
    B("enableSizeAttenuation", pr->enableSizeAttenuation);
    B("enableBrightnessAttenuation", pr->enableBrightnessAttenuation);
    B("enableDOF", pr->enableDOF);
    B("invertPan", pr->invertPan);
    B("invertForwardBack", pr->invertForwardBack);
    B("invertRotation", pr->invertRotation);
    fprintf(f, "    \"effectType\": %d,\n", pr->effectType);
#undef F
#undef B
    fprintf(f, "    \"particleColors\": [");
    for (int i = 0; i < T && i < pr->numColors; i++)
        fprintf(f, "%s\n        {\"r\": %.17g, \"g\": %.17g, \"b\": %.17g}", i ? "," : "",
                (double)pr->particleColors[i].r, (double)pr->particleColors[i].g,
                (double)pr->particleColors[i].b);
    fprintf(f, "\n    ],\n    \"radioByType\": [");
    for (int i = 0; i < T && i < pr->numRadio; i++) fprintf(f, "%s%.17g", i ? ", " : "", (double)pr->radioByType[i]);
    fprintf(f, "],\n    \"rawForceTable\": [");
    for (int i = 0; i < T * T && i < pr->numRawForce; i++)
        fprintf(f, "%s%.17g", i ? ", " : "", (double)pr->rawForceTable[i]);
    fprintf(f, "]\n}\n");
    bool ok = !ferror(f);
    ok = (fclose(f) == 0) && ok;
    return ok ? CF_OK : CF_ERR_IO;
}
