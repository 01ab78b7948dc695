/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the sequential / integer stages of the pymodem
 * demod_chain hot path.  It is the checker that the CUDA path is compared
 * against; nothing in the product path (pymodem_b200/) may import, link or
 * call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it.
 *
 * Parity pin: the reference has no tests or golden vectors of its own
 * (SURVEY.md section 4).  This restatement is pinned against (a) the live
 * Python reference imported in the build container (tools/make_golden.py ->
 * the tests/golden fixtures, checked by tests/test_oracle_golden.py) and (b) the
 * integer known-answer vectors of SURVEY.md Appendix B.2.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference's modems_codecs/).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the slicer clock and
 * the PSK loops must round every double operation separately, exactly like
 * CPython floats do).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* BinarySlicer.slice -- slicer.py:59-107, state from tune() slicer.py:49-56 */
/* ------------------------------------------------------------------------ */
typedef struct {
	double phase_clock;
	double samples_per_symbol;
	double rollover_threshold;
	double lock_rate;
	double last_sample;
	double last_q_sample;          /* quadrature only */
	int64_t streamaddress;
	uint32_t working_byte;
	uint32_t working_bit_count;
	uint32_t state_register;       /* quadrature only */
	uint32_t state_mask;           /* quadrature only */
	uint32_t bits_per_symbol;      /* quadrature only */
	uint32_t demap[16];            /* quadrature only */
} orc_slicer;

ORC_API void orc_slicer_init(orc_slicer *s, double sample_rate, double symbol_rate, double lock_rate)
{
	memset(s, 0, sizeof(*s));
	s->phase_clock = 0.0;
	s->samples_per_symbol = sample_rate / symbol_rate;              /* slicer.py:51 */
	s->rollover_threshold = (s->samples_per_symbol / 2.0) - 0.5;    /* slicer.py:52 */
	s->lock_rate = lock_rate;
	s->last_sample = 0.0;
	s->streamaddress = 0;
}

ORC_API void orc_qslicer_init(orc_slicer *s, double sample_rate, double symbol_rate, double lock_rate,
                              uint32_t state_mask, uint32_t bits_per_symbol, const uint32_t *demap)
{
	orc_slicer_init(s, sample_rate, symbol_rate, lock_rate);        /* slicer.py:181-191 */
	s->state_mask = state_mask;
	s->bits_per_symbol = bits_per_symbol;
	memcpy(s->demap, demap, sizeof(uint32_t) * (state_mask + 1u));
}

/* returns number of AddressedData (byte, address) pairs appended */
ORC_API int64_t orc_binary_slice(orc_slicer *s, const double *samples, int64_t n,
                                 uint8_t *out_bytes, int64_t *out_addr, int64_t cap)
{
	int64_t count = 0;
	for (int64_t i = 0; i < n; i++) {
		double sample = samples[i];
		s->streamaddress += 1;                                      /* :75 */
		s->phase_clock += 1.0;                                      /* :77 */
		if (s->phase_clock >= s->rollover_threshold) {              /* :79 */
			s->phase_clock -= s->samples_per_symbol;                /* :81 */
			s->working_byte = (s->working_byte << 1) & 0xFF;        /* :83 */
			if (sample >= 0) s->working_byte |= 1;                  /* :85-90 */
			else s->working_byte &= 0xFE;
			s->working_bit_count += 1;
			if (s->working_bit_count >= 8) {                        /* :95-97 */
				s->working_bit_count = 0;
				if (count < cap) {
					out_bytes[count] = (uint8_t)s->working_byte;
					out_addr[count] = s->streamaddress;
				}
				count++;
			}
		}
		if ((s->last_sample < 0.0 && sample >= 0.0) ||              /* :99-102 */
		    (s->last_sample >= 0.0 && sample < 0.0)) {
			s->phase_clock = s->phase_clock * s->lock_rate;         /* :104 */
		}
		s->last_sample = sample;                                    /* :106 */
	}
	return count;
}

/* QuadratureSlicer.slice -- slicer.py:193-242 */
ORC_API int64_t orc_quadrature_slice(orc_slicer *s, const double *i_samples, const double *q_samples,
                                     int64_t n, uint8_t *out_bytes, int64_t *out_addr, int64_t cap)
{
	int64_t count = 0;
	for (int64_t k = 0; k < n; k++) {
		double i_sample = i_samples[k], q_sample = q_samples[k];
		s->streamaddress += 1;                                      /* :199 */
		s->phase_clock += 1.0;
		if (s->phase_clock >= s->rollover_threshold) {              /* :203 */
			s->phase_clock -= s->samples_per_symbol;
			s->state_register = (s->state_register << 2) & s->state_mask;  /* :209 */
			if (i_sample >= 0) s->state_register |= 2;
			if (q_sample >= 0) s->state_register |= 1;
			s->working_byte = s->working_byte << s->bits_per_symbol;       /* :215 */
			s->working_byte |= s->demap[s->state_register];
			s->working_bit_count += s->bits_per_symbol;
			if (s->working_bit_count >= 8) {                        /* :221-224 */
				s->working_bit_count = 0;
				s->working_byte &= 0xFF;
				if (count < cap) {
					out_bytes[count] = (uint8_t)s->working_byte;
					out_addr[count] = s->streamaddress;
				}
				count++;
			}
		}
		if (((s->last_sample < 0.0 && i_sample >= 0.0) || (s->last_sample >= 0.0 && i_sample < 0.0)) ||
		    ((s->last_q_sample < 0.0 && q_sample >= 0.0) || (s->last_q_sample >= 0.0 && q_sample < 0.0))) {
			s->phase_clock = s->phase_clock * s->lock_rate;         /* :226-235 */
		}
		s->last_sample = i_sample;
		s->last_q_sample = q_sample;
	}
	return count;
}

/* ------------------------------------------------------------------------ */
/* LFSR.stream_unscramble_8bit -- lfsr.py:22-52 (addresses pass through)      */
/* ------------------------------------------------------------------------ */
typedef struct {
	uint64_t polynomial;
	uint64_t shift_register;
	int invert;
} orc_lfsr;

ORC_API void orc_lfsr_init(orc_lfsr *l, uint64_t poly, int invert)
{
	l->polynomial = poly;
	l->shift_register = 0;                                          /* lfsr.py:16 */
	l->invert = invert;
}

ORC_API void orc_lfsr_unscramble(orc_lfsr *l, const uint8_t *in, uint8_t *out, int64_t n)
{
	uint32_t working_byte = 0;                                      /* :30, persists over bytes */
	for (int64_t k = 0; k < n; k++) {
		uint32_t input_byte = in[k];
		for (int b = 0; b < 8; b++) {
			working_byte <<= 1;
			working_byte &= 0xFE;
			if (input_byte & 0x80) l->shift_register ^= l->polynomial;   /* :38-40 */
			working_byte |= (uint32_t)(l->shift_register & 1);
			input_byte <<= 1;
			l->shift_register >>= 1;                                /* :44 */
		}
		out[k] = l->invert ? (uint8_t)(0xFF ^ working_byte) : (uint8_t)working_byte;
	}
}

/* ------------------------------------------------------------------------ */
/* CheckCRC / AppendCRC -- crc_functions.py:9-76                             */
/* ------------------------------------------------------------------------ */
static uint32_t crc16_x25(const uint8_t *p, int64_t n)
{
	uint32_t crc = 0xFFFF;
	for (int64_t k = 0; k < n; k++) {
		uint32_t byte = p[k];
		for (int i = 0; i < 8; i++) {
			if ((crc & 1) != (byte & 1)) crc = (crc >> 1) ^ 0x8408;  /* :48-51 */
			else crc >>= 1;
			byte >>= 1;
		}
	}
	return crc ^ 0xFFFF;
}

/* out[0]=carried, out[1]=calculated, out[2]=valid.  Needs n >= 2 (the
 * reference indexes packet[-1], packet[-2]; shorter packets never reach it). */
ORC_API void orc_check_crc(const uint8_t *packet, int64_t n, uint32_t *out)
{
	uint32_t carried = (uint32_t)packet[n - 1] * 256u + packet[n - 2];   /* :44 */
	uint32_t calc = crc16_x25(packet, n - 2);
	out[0] = carried;
	out[1] = calc;
	out[2] = (carried == calc);      /* Hamming distance 0 <=> equal, :55-61 */
}

ORC_API uint32_t orc_crc16(const uint8_t *p, int64_t n) { return crc16_x25(p, n); }

/* ValidateHeader -- packet_meta.py:21-41 (subfield_character_index never
 * resets, so only bytes 0..6 are range-checked) */
ORC_API int orc_validate_header(const uint8_t *frame, int64_t count)
{
	int result = 1;
	if (count > 15) {
		for (int64_t index = 0; index < count && index < 7; index++) {
			uint32_t wc = frame[index] >> 1;
			if ((wc < 32 || wc > 126) && wc != 0) result = 0;
		}
	} else {
		result = 0;
	}
	return result;
}

/* ------------------------------------------------------------------------ */
/* AX25Codec.decode -- ax25.py:25-93                                         */
/* The working packet's data list is unbounded in the reference; here it is  */
/* a growable byte vector.                                                   */
/* ------------------------------------------------------------------------ */
typedef struct {
	uint32_t working_byte;
	int64_t byte_index;
	int64_t one_count;
	int64_t bit_index;
	int64_t min_packet_length;
	int64_t max_packet_length;
	uint8_t *data;      /* working_packet.data */
	int64_t len, cap;
} orc_ax25;

ORC_API orc_ax25 *orc_ax25_new(void)
{
	orc_ax25 *a = (orc_ax25 *)calloc(1, sizeof(orc_ax25));
	a->min_packet_length = 18;                                      /* ax25.py:14 */
	a->max_packet_length = 1023;                                    /* ax25.py:15 */
	a->cap = 4096;
	a->data = (uint8_t *)malloc((size_t)a->cap);
	return a;
}

ORC_API void orc_ax25_free(orc_ax25 *a) { free(a->data); free(a); }

static void ax25_append(orc_ax25 *a, uint8_t b)
{
	if (a->len == a->cap) {
		a->cap *= 2;
		a->data = (uint8_t *)realloc(a->data, (size_t)a->cap);
	}
	a->data[a->len++] = b;
}

/*
 * Packet records are written to a flat arena:
 *   rec_addr[r]  streamaddress, rec_off[r] offset into arena, rec_len[r] length.
 * Returns the number of packets emitted by this call; *arena_used is
 * advanced.  Records beyond rec_cap / bytes beyond arena_cap are counted but
 * not stored (caller re-runs with bigger buffers).
 */
ORC_API int64_t orc_ax25_decode(orc_ax25 *a, const uint8_t *bytes, const int64_t *addr, int64_t n,
                                int64_t *rec_addr, int64_t *rec_off, int64_t *rec_len, int64_t rec_cap,
                                uint8_t *arena, int64_t arena_cap, int64_t *arena_used)
{
	int64_t nrec = 0;
	for (int64_t k = 0; k < n; k++) {
		uint32_t input_byte = bytes[k];
		for (int b = 0; b < 8; b++) {
			if (input_byte & 0x80) {                                /* :30 '1' bit */
				a->working_byte |= 0x80;
				a->one_count += 1;
				a->bit_index += 1;
				if (a->one_count > 6) {                             /* :35-38 abort (data NOT cleared) */
					a->bit_index = 0;
					a->byte_index = 0;
				}
				if (a->bit_index == 8) {                            /* :39-51 */
					a->bit_index = 0;
					ax25_append(a, (uint8_t)a->working_byte);
					a->byte_index += 1;
					if (a->byte_index > a->max_packet_length) {
						a->byte_index = 0;
						a->one_count = 0;
					}
				}
				a->working_byte >>= 1;                              /* :52 */
			} else {                                                /* :53 '0' bit */
				if (a->one_count < 5) {
					a->bit_index += 1;
					if (a->bit_index == 8) {
						a->bit_index = 0;
						ax25_append(a, (uint8_t)a->working_byte);
						a->byte_index += 1;
						if (a->byte_index > a->max_packet_length) a->byte_index = 0;   /* :63-68 */
					}
					a->working_byte >>= 1;                          /* :69 */
				} else if (a->one_count == 5) {
					/* stuffed zero ignored :70-72 */
				} else if (a->one_count == 6) {                     /* :73 flag */
					if (a->byte_index >= a->min_packet_length && a->bit_index == 7) {
						if (nrec < rec_cap && *arena_used + a->len <= arena_cap) {
							rec_addr[nrec] = addr[k];               /* :82 */
							rec_off[nrec] = *arena_used;
							rec_len[nrec] = a->len;
							memcpy(arena + *arena_used, a->data, (size_t)a->len);
						}
						*arena_used += a->len;
						nrec++;
					}
					a->len = 0;                                     /* :88 new PacketMeta() */
					a->byte_index = 0;
					a->bit_index = 0;
				}
				a->one_count = 0;                                   /* :91 */
			}
			input_byte <<= 1;
		}
	}
	return nrec;
}
