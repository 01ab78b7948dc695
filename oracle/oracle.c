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

/* ------------------------------------------------------------------------ */
/* GF(2^8) tables -- gf_functions.py:47-74 (initialize), :18-24 (mul)        */
/* ------------------------------------------------------------------------ */
typedef struct {
	int order;               /* 2**power */
	int table[256];          /* antilog, order-1 entries used (gf_functions.py:52-54) */
	int index[256];          /* log; index[0] stays 0 */
	int inverse[256];
} orc_gf;

static int gf_mul(const orc_gf *gf, int a, int b)
{
	if (a == 0 || b == 0) return 0;                                 /* :19-20 */
	int result = gf->index[a] + gf->index[b];
	while (result > gf->order - 2) result -= gf->order - 1;         /* :22-23 */
	return gf->table[result];
}

static void gf_init(orc_gf *gf, int power, int genpoly)
{
	memset(gf, 0, sizeof(*gf));
	gf->order = 1 << power;
	int lfsr = 1;                                                   /* :62 a^0 */
	for (int i = gf->order - 2; i >= 0; i--) {                      /* :63-66 */
		int feedback = lfsr & 1;                                    /* lfsr_step :7-16 */
		lfsr >>= 1;
		if (feedback) lfsr ^= genpoly >> 1;
		gf->table[i] = lfsr;
		gf->index[lfsr] = i;
	}
	for (int i = 1; i < gf->order; i++) {                           /* :70-74 brute-force inverse */
		int j = 1;
		while (gf_mul(gf, i, j) != 1) j++;
		gf->inverse[i] = j;
	}
}

/* ------------------------------------------------------------------------ */
/* Reed-Solomon -- rs_functions.py:9-31 (initialize), :33-150 (decode)       */
/* ------------------------------------------------------------------------ */
typedef struct {
	orc_gf gf;
	int first_root, num_roots;
	int genpoly[64];
} orc_rs;

static void rs_init(orc_rs *rs, int first_root, int num_roots, int gf_power, int gf_poly)
{
	gf_init(&rs->gf, gf_power, gf_poly);
	rs->first_root = first_root;
	rs->num_roots = num_roots;
	memset(rs->genpoly, 0, sizeof(rs->genpoly));
	rs->genpoly[0] = rs->gf.table[first_root];                      /* :19 */
	rs->genpoly[1] = 1;
	int len = 2;
	for (int i = first_root + 1; i < first_root + num_roots; i++) { /* :22-30: multiply by (x + a^i) */
		int f0 = rs->gf.table[i];
		int res[64] = {0};
		for (int a = 0; a < len; a++) {                             /* gf_functions.convolve :36-45 */
			res[a] ^= gf_mul(&rs->gf, rs->genpoly[a], f0);
			res[a + 1] ^= gf_mul(&rs->gf, rs->genpoly[a], 1);
		}
		len += 1;
		memcpy(rs->genpoly, res, sizeof(int) * (size_t)len);
	}
}

/* returns number of corrected bytes, or -1; corrects data in place */
static int rs_decode(const orc_rs *rs, int *data, int block_size, int min_distance)
{
	const orc_gf *gf = &rs->gf;
	const int nr = rs->num_roots;
	int syndromes[32], error_locator[32], next_error_locator[32], error_locations[32], error_magnitudes[32];
	int correction_poly[33];
	int error_count = 0;
	for (int i = 0; i < nr; i++) {                                  /* :36-42 */
		syndromes[i] = 0;
		int x = gf->table[rs->first_root + i];
		for (int j = 0; j < block_size - 1; j++) syndromes[i] = gf_mul(gf, syndromes[i] ^ data[j], x);
		syndromes[i] ^= data[block_size - 1];
	}
	for (int i = 0; i < nr; i++) {                                  /* :50-56 */
		error_locator[i] = 0; next_error_locator[i] = 0; error_locations[i] = 0; error_magnitudes[i] = 0;
	}
	for (int i = 0; i < nr + 1; i++) correction_poly[i] = 0;
	error_locator[0] = 1;
	correction_poly[1] = 1;
	int order_tracker = 0;
	for (int step_factor = 1; step_factor <= nr; step_factor++) {   /* :60-82 Berlekamp */
		int y = step_factor - 1;
		int e = syndromes[y];
		for (int i = 1; i <= order_tracker; i++) {
			int x = y - i;
			e ^= gf_mul(gf, error_locator[i], syndromes[x]);
		}
		if (e != 0) {
			for (int i = 0; i <= order_tracker; i++)
				next_error_locator[i] = error_locator[i] ^ gf_mul(gf, e, correction_poly[i]);
			e = gf->inverse[e];
			for (int i = 0; i < nr / 2 + 1; i++) correction_poly[i] = gf_mul(gf, error_locator[i], e);
			for (int i = 0; i < nr / 2 + 1; i++) error_locator[i] = next_error_locator[i];
		}
		if (2 * order_tracker < step_factor) order_tracker = step_factor - order_tracker;
		for (int i = nr; i > 0; i--) correction_poly[i] = correction_poly[i - 1];
		correction_poly[0] = 0;
	}
	for (int j = 0; j < block_size; j++) {                          /* :85-98 Chien search */
		int x = 0;
		int y = j + gf->order - block_size;
		for (int i = 1; i < nr / 2 + 1; i++) {
			if (error_locator[i]) {
				int z = y * i + gf->index[error_locator[i]];
				while (z > gf->order - 2) z -= gf->order - 1;
				x ^= gf->table[z];
			}
		}
		x ^= error_locator[0];
		if (x == 0) {
			error_locations[error_count] = j;
			error_count++;
		}
	}
	if (error_count <= nr / 2 - min_distance) {                     /* :99 */
		for (int i = 0; i < error_count; i++) {                     /* :101-108 Forney */
			correction_poly[i] = syndromes[rs->first_root + i];
			for (int j = 1; j <= i; j++)
				correction_poly[i] ^= gf_mul(gf, syndromes[rs->first_root + i - j], error_locator[j]);
		}
		for (int i = 0; i < error_count; i++) {
			int e = block_size - error_locations[i] - 1;
			int z = correction_poly[0];
			for (int j = 1; j < error_count; j++) {                 /* :112-122 */
				int x = e * j;
				while (x > gf->order - 2) x -= gf->order - 1;
				x = gf->order - x - 1;
				while (x > gf->order - 2) x -= gf->order - 1;
				z ^= gf_mul(gf, correction_poly[j], gf->table[x]);
			}
			z = gf_mul(gf, z, gf->table[e]);
			int y = error_locator[1];
			for (int j = 3; j < nr / 2 + 1; j += 2) {               /* :125-132 */
				int x = e * (j - 1);
				while (x > gf->order - 2) x -= gf->order - 1;
				x = gf->order - x - 1;
				while (x > gf->order - 2) x -= gf->order - 1;
				y ^= gf_mul(gf, error_locator[j], gf->table[x]);
			}
			y = gf->index[y];
			y = gf->order - y - 1;
			if (y == gf->order - 1) y = 0;
			y = gf->table[y];
			error_magnitudes[i] = gf_mul(gf, y, z);
			data[error_locations[i]] ^= error_magnitudes[i];
		}
	}
	for (int i = 0; i < nr; i++) {                                  /* :142-149 re-check */
		syndromes[i] = 0;
		int x = gf->table[rs->first_root + i];
		for (int j = 0; j < block_size - 1; j++) syndromes[i] = gf_mul(gf, syndromes[i] ^ data[j], x);
		syndromes[i] ^= data[block_size - 1];
		if (syndromes[i] != 0) return -1;
	}
	return error_count;
}

/* KAT access for the tests */
ORC_API void orc_gf_tables(int *table, int *index, int *inverse)
{
	orc_gf gf;
	gf_init(&gf, 8, 0x11D);
	memcpy(table, gf.table, sizeof(int) * 255);
	memcpy(index, gf.index, sizeof(int) * 256);
	memcpy(inverse, gf.inverse, sizeof(int) * 256);
}

ORC_API int orc_rs_genpoly(int num_roots, int *out)
{
	orc_rs rs;
	rs_init(&rs, 0, num_roots, 8, 0x11D);
	memcpy(out, rs.genpoly, sizeof(int) * (size_t)(num_roots + 1));
	return num_roots + 1;
}

ORC_API int orc_rs_decode(int num_roots, uint8_t *data, int block_size, int min_distance)
{
	orc_rs rs;
	int buf[256];
	rs_init(&rs, 0, num_roots, 8, 0x11D);
	for (int i = 0; i < block_size; i++) buf[i] = data[i];
	int r = rs_decode(&rs, buf, block_size, min_distance);
	for (int i = 0; i < block_size; i++) data[i] = (uint8_t)buf[i];
	return r;
}

/* ------------------------------------------------------------------------ */
/* IL2P -- il2p.py                                                           */
/* ------------------------------------------------------------------------ */
static const uint8_t hamming_decode_table[128] = {                  /* il2p.py:19-42 */
	0x0, 0x0, 0x0, 0x3, 0x0, 0x5, 0xe, 0x7, 0x0, 0x9, 0xe, 0xb, 0xe, 0xd, 0xe, 0xe,
	0x0, 0x3, 0x3, 0x3, 0x4, 0xd, 0x6, 0x3, 0x8, 0xd, 0xa, 0x3, 0xd, 0xd, 0xe, 0xd,
	0x0, 0x5, 0x2, 0xb, 0x5, 0x5, 0x6, 0x5, 0x8, 0xb, 0xb, 0xb, 0xc, 0x5, 0xe, 0xb,
	0x8, 0x1, 0x6, 0x3, 0x6, 0x5, 0x6, 0x6, 0x8, 0x8, 0x8, 0xb, 0x8, 0xd, 0x6, 0xf,
	0x0, 0x9, 0x2, 0x7, 0x4, 0x7, 0x7, 0x7, 0x9, 0x9, 0xa, 0x9, 0xc, 0x9, 0xe, 0x7,
	0x4, 0x1, 0xa, 0x3, 0x4, 0x4, 0x4, 0x7, 0xa, 0x9, 0xa, 0xa, 0x4, 0xd, 0xa, 0xf,
	0x2, 0x1, 0x2, 0x2, 0xc, 0x5, 0x2, 0x7, 0xc, 0x9, 0x2, 0xb, 0xc, 0xc, 0xc, 0xf,
	0x1, 0x1, 0x2, 0x1, 0x4, 0x1, 0x6, 0xf, 0x8, 0x1, 0xa, 0xf, 0xc, 0xf, 0xf, 0xf
};

static int bit_distance_32(uint32_t a, uint32_t b)                 /* il2p.py:44-86 == popcount(a ^ b) */
{
	return __builtin_popcount(a ^ b);
}

typedef struct {
	/* options (il2p.py:111-116, 140-145) */
	int collect_trailing_crc, min_distance, disable_rs, sync_tolerance;
	/* state (il2p.py:118-138) */
	int state;                      /* 0 sync_search, 1 rx_header, 2 rx_bigblocks, 3 rx_smallblocks, 4 rx_trailing_crc */
	uint32_t working_word;
	int buffer[255];
	int bit_index, byte_index_a, block_index;
	int bytes_corrected, block_fail;
	int block_count, block_size, big_blocks;
	int count_subfield;
	orc_rs header_rs, block_rs;
	uint8_t *data;                  /* working_packet.data */
	int64_t len, cap;
} orc_il2p;

ORC_API orc_il2p *orc_il2p_new(int crc, int disable_rs, int min_dist, int sync_tol)
{
	orc_il2p *c = (orc_il2p *)calloc(1, sizeof(orc_il2p));
	c->collect_trailing_crc = crc;
	c->disable_rs = disable_rs;
	c->min_distance = min_dist;
	c->sync_tolerance = sync_tol;
	c->state = 0;
	c->working_word = 0xFFFFFF;                                     /* :119 */
	rs_init(&c->header_rs, 0, 2, 8, 0x11D);                         /* :130-135 */
	rs_init(&c->block_rs, 0, 16, 8, 0x11D);
	c->cap = 4096;
	c->data = (uint8_t *)malloc((size_t)c->cap);
	return c;
}

ORC_API void orc_il2p_free(orc_il2p *c) { free(c->data); free(c); }

static void il2p_append(orc_il2p *c, int b)
{
	if (c->len == c->cap) {
		c->cap *= 2;
		c->data = (uint8_t *)realloc(c->data, (size_t)c->cap);
	}
	c->data[c->len++] = (uint8_t)b;
}

/* block_unscramble :160-163 with LFSRnoaddr.stream_unscramble_8bit lfsr.py:62-92 (poly 0x211, no invert) */
static void il2p_block_unscramble(orc_il2p *c)
{
	uint32_t sr = 0x1F0;
	uint32_t working_byte = 0;
	for (int k = 0; k < c->byte_index_a; k++) {
		uint32_t input_byte = (uint32_t)c->buffer[k];
		for (int b = 0; b < 8; b++) {
			working_byte <<= 1;
			working_byte &= 0xFE;
			if (input_byte & 0x80) sr ^= 0x211;
			working_byte |= sr & 1;
			input_byte <<= 1;
			sr >>= 1;
		}
		c->buffer[k] = (int)working_byte;
	}
}

static void il2p_rs(orc_il2p *c, const orc_rs *rs)                  /* :170-207 */
{
	int r = c->disable_rs ? 0 : rs_decode(rs, c->buffer, c->byte_index_a, c->min_distance);
	if (r < 0) c->block_fail = 1;
	else c->bytes_corrected += r;
}

static void il2p_append_crc(orc_il2p *c)                            /* crc_functions.py:63-76 */
{
	uint32_t crc = crc16_x25(c->data, c->len);
	il2p_append(c, (int)(crc & 0xFF));
	il2p_append(c, (int)(crc >> 8));
}

/* unpack_il2p_header :214-290 + construct_ax25_header :292-344 + reform_control_byte :89-107 */
static void il2p_header_to_ax25(orc_il2p *c)
{
	const int *b = c->buffer;
	int type_subfield = (b[1] & 0x80) >> 7;
	int count = 0, pid = 0, control = 0;
	for (int i = 0; i < 10; i++) if (b[i + 2] & 0x80) count |= 0x200 >> i;
	for (int i = 0; i < 4; i++) if (b[i + 1] & 0x40) pid |= 0x8 >> i;
	for (int i = 0; i < 7; i++) if (b[i + 5] & 0x40) control |= 0x40 >> i;
	c->count_subfield = count;
	int dest[7], source[7];
	for (int i = 0; i < 6; i++) dest[i] = (b[i] & 0x3F) + 0x20;
	dest[6] = b[12] >> 4;
	for (int i = 0; i < 6; i++) source[i] = (b[i + 6] & 0x3F) + 0x20;
	source[6] = b[12] & 0xF;
	enum { T_UI, T_S, T_U, T_I } type;
	if (b[0] & 0x40) type = T_UI;
	else if (pid == 0x0) type = T_S;
	else if (pid == 0x1) type = T_U;
	else type = T_I;
	static const int pid_table[16] = {0, 0, 0x10, 0x01, 0x06, 0x07, 0x08, 0xC3, 0xC4, 0xCA, 0xCB, 0xCC, 0xCD, 0xCE, 0xCF, 0xF0};
	int pid_byte = pid_table[pid];
	int pf = (control & 0x40) != 0, cbit = 0, nr = 0, ns = 0, opcode = 0;
	if (type == T_I) { ns = control & 0x7; nr = (control >> 3) & 0x7; cbit = 1; }
	else if (type == T_S) { nr = (control >> 3) & 0x7; if (control & 0x4) cbit = 1; opcode = control & 0x3; }
	else { if (control & 0x4) cbit = 1; opcode = (control >> 3) & 0x7; }
	if (type_subfield != 1) return;                                 /* transparent: nothing reconstructed (:342) */
	for (int i = 0; i < 6; i++) il2p_append(c, dest[i] << 1);
	int v = (dest[6] << 1) + 0x60;
	if (cbit) v += 0x80;
	il2p_append(c, v);
	for (int i = 0; i < 6; i++) il2p_append(c, source[i] << 1);
	v = (source[6] << 1) + 0x60;
	if (!cbit) v += 0x80;
	v += 1;
	il2p_append(c, v);
	static const int u_control[8] = {0x2F, 0x43, 0x0F, 0x63, 0x87, 0x03, 0xAF, 0xE3};
	int cb = 0;
	if (type == T_U || type == T_UI) { cb = u_control[opcode]; if (pf) cb |= 0x10; }
	else if (type == T_S) { cb = 0x1 | (opcode << 2) | (nr << 5); if (pf) cb |= 0x10; }
	else { cb = (ns << 1) | (nr << 5); if (pf) cb |= 0x10; }
	il2p_append(c, cb);
	if (pid_byte != 0) il2p_append(c, pid_byte);
}

/*
 * IL2PCodec.decode -- il2p.py:360-519.  Records as in orc_ax25_decode plus rec_corr[]
 * (PacketMeta.BytesCorrected).  Every value the reference appends to working_packet.data fits a
 * byte (callsign characters <= 0x5F << 1, SSID bytes <= 30 + 0x60 + 0x80 + 1).
 */
ORC_API int64_t orc_il2p_decode(orc_il2p *c, const uint8_t *bytes, const int64_t *addr, int64_t n,
                                int64_t *rec_addr, int64_t *rec_off, int64_t *rec_len, int64_t *rec_corr,
                                int64_t rec_cap, uint8_t *arena, int64_t arena_cap, int64_t *arena_used)
{
	int64_t nrec = 0;
#define WRITE_N_SEARCH()                                                          \
	do {                                                                          \
		if (nrec < rec_cap && *arena_used + c->len <= arena_cap) {                \
			rec_addr[nrec] = addr[k]; rec_off[nrec] = *arena_used;                \
			rec_len[nrec] = c->len; rec_corr[nrec] = c->bytes_corrected;          \
			memcpy(arena + *arena_used, c->data, (size_t)c->len);                 \
		}                                                                         \
		*arena_used += c->len; nrec++;                                            \
		c->bytes_corrected = 0; c->len = 0; c->state = 0;                         \
	} while (0)
	for (int64_t k = 0; k < n; k++) {
		uint32_t input_byte = bytes[k];
		for (int ib = 0; ib < 8; ib++) {
			uint32_t mask = (c->state == 0) ? 0xFFFFFFFFu : 0xFFu;  /* get_a_bit :147-153 */
			c->working_word = (c->working_word << 1) & mask;
			if (input_byte & 0x80) c->working_word |= 1;
			input_byte <<= 1;
			c->bit_index += 1;
			if (c->state == 0) {                                    /* :367-376 */
				if (bit_distance_32(c->working_word & 0xFFFFFF, 0xF15E48) <= c->sync_tolerance ||
				    bit_distance_32(c->working_word, 0x5D57DF7F) <= c->sync_tolerance) {
					c->bit_index = 0;
					c->state = 1;
				}
				continue;
			}
			if (c->bit_index != 8) continue;
			c->bit_index = 0;
			c->buffer[c->byte_index_a] = (int)c->working_word;
			c->byte_index_a += 1;
			if (c->state == 1) {                                    /* rx_header :377-434 */
				if (c->byte_index_a != 15) continue;
				il2p_rs(c, &c->header_rs);
				c->byte_index_a = 13;
				il2p_block_unscramble(c);
				c->byte_index_a = 0;
				c->block_index = 0;
				il2p_header_to_ax25(c);
				if (c->block_fail) {
					c->block_fail = 0;
					c->state = 0;
					c->len = 0;
				} else if (c->count_subfield > 0) {
					int cnt = c->count_subfield;                    /* calc_big_small_blocks :346-358 */
					c->block_count = (cnt + 238) / 239;
					c->block_size = cnt / c->block_count;
					c->big_blocks = cnt - c->block_count * c->block_size;
					if (c->big_blocks > 0) { c->block_size += 1; c->state = 2; }
					else c->state = 3;
					c->bit_index = 0;
				} else if (c->collect_trailing_crc) {
					c->state = 4;
				} else {
					il2p_append_crc(c);
					WRITE_N_SEARCH();
				}
			} else if (c->state == 2 || c->state == 3) {            /* :436-501 */
				if (c->byte_index_a != c->block_size + 16) continue;
				il2p_rs(c, &c->block_rs);
				il2p_block_unscramble(c);
				for (int i = 0; i < c->block_size; i++) il2p_append(c, c->buffer[i]);
				c->block_index += 1;
				c->byte_index_a = 0;
				if (c->block_fail) {
					c->block_fail = 0;
					c->len = 0;
					c->state = 0;
				} else if (c->state == 2 && c->block_index == c->big_blocks) {
					if (c->block_count > c->block_index) { c->block_size -= 1; c->state = 3; }
					else if (c->collect_trailing_crc) c->state = 4;
					else { il2p_append_crc(c); WRITE_N_SEARCH(); }
				} else if (c->state == 3 && c->block_index == c->block_count) {
					if (c->collect_trailing_crc) c->state = 4;
					else { il2p_append_crc(c); WRITE_N_SEARCH(); }
				}
			} else if (c->state == 4) {                             /* rx_trailing_crc :502-518 */
				if (c->byte_index_a != 4) continue;
				c->byte_index_a = 0;
				int trailing_crc = 0;
				for (int i = 0; i < 4; i++) trailing_crc += hamming_decode_table[c->buffer[i] & 0x7F] << (12 - i * 4);
				il2p_append(c, trailing_crc & 0xFF);
				il2p_append(c, trailing_crc >> 8);
				WRITE_N_SEARCH();
			}
		}
	}
#undef WRITE_N_SEARCH
	return nrec;
}

/* ------------------------------------------------------------------------ */
/* PSK / PLL recursive loops: agc.py, nco.py, iir.py, pi_control.py,          */
/* phase_detector.py, complexmath.py, psk.py:162-195 / 705-773,              */
/* afsk_pll.py:140-170.  Every double operation rounds separately            */
/* (-ffp-contract=off), in the reference's evaluation order.                 */
/* ------------------------------------------------------------------------ */
typedef struct {
	/* AGC -- agc.py:8-24 */
	double agc_scaled_attack;      /* attack_rate / sample_rate          agc.py:15 */
	double agc_scaled_decay;       /* decay_rate / sample_rate           agc.py:16 */
	double agc_sustain_time;
	double agc_sustain_increment;  /* 1 / sample_rate                    agc.py:17 */
	double agc_target;
	/* NCO -- nco.py:14-32 */
	double nco_phase_scale;        /* 2.0 * pi / sample_rate             nco.py:30 */
	double nco_index_scale;        /* wavetable_size / (2.0 * pi)        nco.py:28 */
	double nco_set_frequency;
	double nco_two_pi;             /* 2.0 * pi */
	double nco_quarter;            /* wavetable_size / 4.0               nco.py:46 */
	const double *nco_wavetable;   /* amplitude * sin(i * 2.0 * pi / size), built by the caller with math.sin */
	int64_t nco_size;
	/* IIR_1 -- iir.py:15-29 */
	double iir_b0, iir_b1, iir_a1;
	/* PI_control -- pi_control.py:8-13 */
	double pi_gain, pi_p, pi_i, pi_limit, pi_integral0;
	/* PhaseDetector.qpsk_error_table -- phase_detector.py:34-45, granularity x granularity ints */
	const int32_t *pd_table;
	int64_t pd_granularity;
} orc_loop;

typedef struct {
	double envelope, sustain_count, normal;                         /* agc.py:18-24 */
	double phase, control, sine, cosine;                            /* nco.py:22-26 */
	double x1, y1;                                                  /* iir.py:33-34 */
	double integral, proportional;                                  /* pi_control.py:12-13 */
} orc_loop_state;

static void loop_state_init(const orc_loop *L, orc_loop_state *s, double normal)
{
	memset(s, 0, sizeof(*s));
	s->normal = normal;
	s->integral = L->pi_integral0;
}

/* AGC.peak_detect + the scaling line of AGC.apply -- agc.py:26-37, 72-76 */
static double agc_step(const orc_loop *L, orc_loop_state *s, double sample)
{
	double compare_value = fabs(sample);
	if (compare_value > s->envelope) {
		s->envelope += (L->agc_scaled_attack * s->normal);
		if (s->envelope > compare_value) s->envelope = compare_value;
		s->sustain_count = 0.0;
	}
	if (s->sustain_count >= L->agc_sustain_time) {
		s->envelope -= (L->agc_scaled_decay * s->normal);
		if (s->envelope < 0) s->envelope = 0;
	}
	s->sustain_count += L->agc_sustain_increment;
	if (s->envelope != 0) return L->agc_target * sample / s->envelope;
	return sample;
}

/* NCO.update -- nco.py:34-53 */
static void nco_step(const orc_loop *L, orc_loop_state *s)
{
	s->phase += (L->nco_phase_scale * (L->nco_set_frequency + s->control));
	while (s->phase >= L->nco_two_pi) s->phase = s->phase - L->nco_two_pi;
	while (s->phase < 0) s->phase = s->phase + L->nco_two_pi;
	int64_t sine_index = (int64_t)(s->phase * L->nco_index_scale);
	if (sine_index < L->nco_size) s->sine = L->nco_wavetable[sine_index];   /* IndexError keeps the old value :41-45 */
	int64_t cosine_index = (int64_t)((double)sine_index + L->nco_quarter);
	while (cosine_index >= L->nco_size) cosine_index -= L->nco_size;
	while (cosine_index < 0) cosine_index += L->nco_size;
	s->cosine = L->nco_wavetable[cosine_index];
}

/* IIR_1.update -- iir.py:38-54 (order 1) */
static double iir_step(const orc_loop *L, orc_loop_state *s, double sample)
{
	double v = 0;
	v += (sample * L->iir_b0);
	v += (s->x1 * L->iir_b1);
	v += (s->y1 * L->iir_a1);
	s->x1 = sample;
	s->y1 = v;
	return v;
}

/* PI_control.update_saturate -- pi_control.py:25-33 */
static double pi_step(const orc_loop *L, orc_loop_state *s, double sample)
{
	s->proportional = L->pi_gain * L->pi_p * sample;
	s->integral += L->pi_gain * (L->pi_i * sample);
	if (s->integral > L->pi_limit) s->integral = L->pi_limit;
	if (s->integral < -L->pi_limit) s->integral = -L->pi_limit;
	return s->proportional + s->integral;
}

/* max(buffer) -- agc.py:67 (python max over a float64 ndarray: first maximal element, NaN-free input) */
ORC_API double orc_buffer_max(const double *x, int64_t n)
{
	double m = x[0];
	for (int64_t i = 1; i < n; i++) if (x[i] > m) m = x[i];
	return m;
}

/* AGC.apply -- agc.py:61-80, in place */
ORC_API void orc_agc_apply(const orc_loop *L, double *x, int64_t n)
{
	orc_loop_state s;
	if (n <= 0) return;
	loop_state_init(L, &s, orc_buffer_max(x, n));
	for (int64_t i = 0; i < n; i++) x[i] = agc_step(L, &s, x[i]);
}

/* BPSKModem.demod, the Costas loop -- psk.py:170-189; x is the AGC'd buffer; out = i_mixer */
ORC_API void orc_bpsk_loop(const orc_loop *L, const double *x, int64_t n, double *out)
{
	orc_loop_state s;
	loop_state_init(L, &s, 0.0);
	for (int64_t k = 0; k < n; k++) {
		double sample = x[k];
		nco_step(L, &s);
		double i_mixer = sample * s.cosine;                         /* ComplexOutput.real = cosine  nco.py:52 */
		double q_mixer = sample * (-s.sine);                        /* ComplexOutput.imag = -sine   nco.py:53 */
		double loop_mixer = i_mixer * q_mixer;
		double f = iir_step(L, &s, loop_mixer);
		s.control = pi_step(L, &s, f);
		out[k] = i_mixer;
	}
}

/* AFSKPLLModem.demod, the PLL -- afsk_pll.py:152-165; out = PI proportional term */
ORC_API void orc_pll_loop(const orc_loop *L, const double *x, int64_t n, double *out)
{
	orc_loop_state s;
	loop_state_init(L, &s, 0.0);
	for (int64_t k = 0; k < n; k++) {
		nco_step(L, &s);
		double mixer = x[k] * s.sine;
		double f = iir_step(L, &s, mixer);
		s.control = pi_step(L, &s, f);
		out[k] = s.proportional;
	}
}

/* PhaseDetector.get_qpsk_angle_error -- phase_detector.py:124-149 */
static int32_t pd_qpsk_error(const orc_loop *L, double re, double im)
{
	const int64_t g = L->pd_granularity;
	double fr = floor(re * (double)g * 0.5), fi = floor(im * (double)g * 0.5);
	if (fr > 1e9) fr = 1e9;
	if (fr < -1e9) fr = -1e9;
	if (fi > 1e9) fi = 1e9;
	if (fi < -1e9) fi = -1e9;
	int64_t real = (int64_t)fr, imag = (int64_t)fi;
	if (real >= g) real = g - 1;
	if (imag >= g) imag = g - 1;
	if (real <= -g) real = -(g - 1);
	if (imag <= -g) imag = -(g - 1);
	if (real >= 0) {
		if (imag >= 0) return L->pd_table[real * g + imag];
		return L->pd_table[(-imag) * g + real];
	}
	if (imag >= 0) return L->pd_table[imag * g + (-real)];
	return L->pd_table[(-real) * g + (-imag)];
}

/* MPSKModem.demod, the decision-directed loop -- psk.py:733-746; outputs the rotated I/Q samples */
ORC_API void orc_mpsk_loop(const orc_loop *L, const double *re, const double *im, int64_t n,
                           double *out_i, double *out_q)
{
	orc_loop_state s;
	loop_state_init(L, &s, 0.0);
	for (int64_t k = 0; k < n; k++) {
		nco_step(L, &s);
		const double c_re = s.cosine, c_im = -s.sine;
		/* ComplexNumber.multiply -- complexmath.py:15-19 */
		double real = (re[k] * c_re) - (im[k] * c_im);
		double imag = (c_re * im[k]) + (re[k] * c_im);
		double f = iir_step(L, &s, (double)pd_qpsk_error(L, real, imag));
		s.control = nearbyint(pi_step(L, &s, f));                   /* python round(): half to even */
		out_i[k] = real;
		out_q[k] = imag;
	}
}
